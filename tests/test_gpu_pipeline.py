"""GPU: the stream-overlapped host-to-host front end returns exactly what the plain call returns,
step after step, with slot reuse (depth 2, 5 steps, different inputs each step)."""
import numpy as np
import pytest
import torch

import naf_b200

pytestmark = pytest.mark.gpu


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def test_host_pipeline_matches_direct_calls():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev)
    pipe = naf_b200.HostPipeline(m, depth=2)
    r = 4
    reduce = lambda out: out[:, :, r // 2::r, r // 2::r]
    got, want = [], []
    for i in range(5):
        img = rnd(10 + i, 2, 3, 32, 32).pin_memory()
        ft = rnd(20 + i, 2, 16, 8, 8).pin_memory()
        res_h, done = pipe.step(img, ft, (32, 32), reduce=reduce)
        done.synchronize()
        got.append(res_h.clone())
        with torch.no_grad():
            want.append(reduce(m(img.to(dev), ft.to(dev), (32, 32))).cpu())
    pipe.drain()
    torch.cuda.synchronize()
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert torch.equal(g, w)
    # without synchronising between steps (what a throughput loop does): last results still right
    outs = []
    for i in range(4):
        img = rnd(30 + i, 2, 3, 32, 32).pin_memory()
        ft = rnd(40 + i, 2, 16, 8, 8).pin_memory()
        outs.append((pipe.step(img, ft, (32, 32), reduce=reduce), img, ft))
    (res_h, done), img, ft = outs[-1]
    done.synchronize()
    with torch.no_grad():
        w = reduce(m(img.to(dev), ft.to(dev), (32, 32))).cpu()
    assert torch.equal(res_h, w)


def test_graphed_forward_matches_eager():
    dev = torch.device("cuda", 0)
    torch.manual_seed(1)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev)
    fast = naf_b200.GraphedNAF(m)
    for shape in [((1, 3, 64, 64), (1, 32, 8, 8), (64, 64)), ((2, 3, 32, 48), (2, 16, 8, 8), (64, 96))]:
        for i in range(3):
            img = rnd(50 + i, *shape[0]).to(dev)
            ft = rnd(60 + i, *shape[1]).to(dev)
            got = fast(img, ft, shape[2]).clone()
            with torch.no_grad():
                want = m(img, ft, shape[2])
            assert got.shape == want.shape and got.stride() == want.stride()
            assert torch.equal(got, want)
    assert len(fast._entries) == 2


def test_graph_replay_survives_other_shapes_and_weight_updates():
    """A, B, A replay with eager calls of other shapes in between (the single-entry RoPE / tap / weight
    caches get replaced: the graphs must keep their own tensors alive), a non-integer-ratio shape (tap
    tables on the device), `empty_cache()`, then an in-place weight update (new graph, right result)."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(2)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev)
    fast = naf_b200.GraphedNAF(m)
    A = ((1, 3, 64, 64), (1, 32, 8, 8), (64, 64))
    Bs = ((2, 3, 32, 48), (2, 16, 8, 8), (64, 96))
    Cn = ((1, 3, 60, 60), (1, 16, 8, 8), (60, 60))     # 60 / 8: non-integer ratio -> tap tables
    def check(shape, seed):
        img, ft = rnd(seed, *shape[0]).to(dev), rnd(seed + 1, *shape[1]).to(dev)
        got = fast(img, ft, shape[2]).clone()
        with torch.no_grad():
            want = m(img, ft, shape[2])
        assert torch.equal(got, want), shape
    check(A, 70)
    check(Bs, 72)
    check(Cn, 74)
    with torch.no_grad():      # eager calls of yet other shapes: every cache now holds something else
        m(rnd(80, 1, 3, 40, 40).to(dev), rnd(81, 1, 8, 10, 10).to(dev), (40, 40))
        m(rnd(82, 1, 3, 33, 33).to(dev), rnd(83, 1, 8, 9, 9).to(dev), (33, 33))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    junk = torch.full((64 << 20,), float("nan"), device=dev)   # re-use whatever was freed
    check(A, 76)
    check(Cn, 78)
    check(Bs, 84)
    del junk
    n_before = len(fast._entries)
    assert n_before == 3
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    check(A, 86)               # weights changed in place -> fresh capture, stale graphs dropped
    assert len(fast._entries) == 1


def test_host_pipeline_full_result_download():
    """Without `reduce` the whole result crosses to the host every step (bench.py's e2e), through one
    pinned buffer or through the pinned ring."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev)
    img, ft = rnd(90, 2, 3, 32, 32).pin_memory(), rnd(91, 2, 16, 8, 8).pin_memory()
    with torch.no_grad():
        want = m(img.to(dev), ft.to(dev), (32, 32)).permute(0, 2, 3, 1).reshape(-1).cpu()
    pipe = naf_b200.HostPipeline(m, depth=2)
    for _ in range(3):
        res_h, done = pipe.step(img, ft, (32, 32))
    done.synchronize()
    assert torch.equal(res_h, want)
    ring = naf_b200.HostPipeline(m, depth=2, host_ring_bytes=4096)
    for _ in range(2):
        res_h, done = ring.step(img, ft, (32, 32))
    done.synchronize()
    n = 4096 // 4
    chunks = want.numel() // n            # 32768 / 1024 = 32 chunks, even count: last one lands in buffer 1
    assert want.numel() % n == 0 and chunks % 2 == 0
    assert torch.equal(res_h, want[(chunks - 2) * n:(chunks - 1) * n])
