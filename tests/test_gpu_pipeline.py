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
