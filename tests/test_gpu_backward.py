"""GPU: the backward kernels (naf_xattn_bwd_f32, naf_rope_kpool_bwd_f32) and the autograd bindings
against (1) gradient fixtures produced by autograd through the UNMODIFIED reference modules
(tests/golden/bwd_*.npz), (2) the fp64 oracle on seeded inputs (gradcheck-style: the same loss, the
same upstream gradient), for every backward kernel able to run the case.

Tolerance: fp32 SIMT kernels with atomics (summation order varies) and the tensor-core cell kernel (three-pass
split-fp16 GEMMs, fp32 accumulation): 3e-5 of the gradient's max |value| against fp64 / the reference's fp32."""
import numpy as np
import pytest
import torch

import golden_util as G
import naf_b200
from naf_b200 import _lib, ops
from oracle import naf_oracle as O

pytestmark = pytest.mark.gpu
REL = 3e-5


def dev():
    return torch.device("cuda", 0)


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def close(got, want, what, rel=REL):
    want = want.to(torch.float64)
    err = (got.detach().cpu().to(torch.float64) - want).abs().max().item()
    scale = max(want.abs().max().item(), 1e-6)
    assert err <= rel * scale + 1e-7, (what, err, scale)


def bwd_algos(q, k, v, dout, heads, K):
    """Every backward kernel able to run the case: generic always; the fp32 cell kernel and the tensor-core cell
    kernel where they accept it (asked for explicitly, so AUTO's choice cannot hide one of them)."""
    algos = [_lib.ALGO_GENERIC, _lib.ALGO_AUTO]
    for algo in (_lib.ALGO_CELL_SIMT, _lib.ALGO_CELL_TC):
        try:
            ops.xattn_bwd(q, k, v, dout, heads, K, algo=algo)
            algos.append(algo)
        except NotImplementedError:
            pass
    return algos


@pytest.mark.parametrize("name", G.names("bwd_xattn_"))
def test_operator_gradients_match_reference_golden(name):
    c = G.bwd_attention_case(name)
    q, k, v, dout = (c[n].to(dev()) for n in ("q", "k", "v", "dout"))
    for algo in bwd_algos(q, k, v, dout, c["heads"], c["K"]):
        dq, dk, dv = ops.xattn_bwd(q, k, v, dout, c["heads"], c["K"], algo=algo)
        close(dq, c["dq"], (name, algo, "dq"))
        close(dk, c["dk"], (name, algo, "dk"))
        close(dv, c["dv"], (name, algo, "dv"))
    # through the module + autograd, like a training loop would
    mod = naf_b200.CrossAttention(dim=q.shape[1], num_heads=c["heads"], kernel_size=(c["K"], c["K"]))
    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = mod(qg, kg, vg, None)
    assert out.requires_grad
    assert (out.detach().cpu() - c["out"]).abs().max().item() <= 1e-3
    out.backward(dout)
    close(qg.grad, c["dq"], (name, "autograd dq"))
    close(kg.grad, c["dk"], (name, "autograd dk"))
    close(vg.grad, c["dv"], (name, "autograd dv"))


@pytest.mark.parametrize("name", G.names("bwd_naf_"))
def test_module_gradients_match_reference_golden(name):
    """NAF.forward + backward (what train.py:127,136 runs): conv encoder under torch autograd in strict
    fp32, attention path through our forward and backward kernels."""
    c = G.bwd_module_case(name)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = naf_b200.NAF(dim=128, kernel_size=7).eval()
        m.load_state_dict(G.module_state())
        m = m.to(dev())
        img = c["image"].to(dev()).requires_grad_(True)
        ft = c["features"].to(dev()).requires_grad_(True)
        out = m(img, ft, c["output_size"])
        assert (out.detach().cpu() - c["out"]).abs().max().item() <= 1e-4
        out.backward(c["dout"].to(dev()))
        close(ft.grad, c["dfeatures"], (name, "dfeatures"), rel=1e-4)
        close(img.grad, c["dimage"], (name, "dimage"), rel=2e-4)
        named = dict(m.named_parameters())
        checked = 0
        for key, want in c["grads"].items():
            if key.endswith("__sums"):
                pname = key[:-len("__sums")].replace("__", ".")
                g = named[pname].grad.double()
                assert abs(g.sum().item() - want[0].item()) <= 2e-4 * max(1.0, want[1].item()), key
                assert abs(g.abs().sum().item() - want[1].item()) <= 2e-4 * max(1.0, want[1].item()), key
            elif key.endswith("__head"):
                pname = key[:-len("__head")].replace("__", ".")
                g = named[pname].grad.reshape(-1)[:want.numel()]
                close(g, want, (name, pname), rel=3e-4)
            else:
                close(named[key.replace("__", ".")].grad, want, (name, key), rel=3e-4)
            checked += 1
        assert checked >= 36
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


GRAD_CASES = [
    # B, D, n_attn, n_rope, C, Hs, Ws, rep, h, w, K
    (1, 256, 4, 4, 128, 56, 56, (1, 1), 4, 4, 3),      # r=14: cell kernel
    (2, 256, 4, 4, 96, 36, 48, (1, 1), 9, 12, 7),      # r=4: 16-pixel cells, non-square, batch 2
    (1, 256, 4, 4, 64, 24, 24, (2, 2), 6, 6, 5),       # replicated guidance (48 <- 24), r=8
    (2, 256, 4, 4, 16, 32, 32, (1, 1), 13, 13, 9),     # training-style non-integer ratio -> generic
    (1, 256, 4, 1, 32, 30, 30, (1, 1), 10, 10, 3),     # rope heads != attention heads: composed operators
    (1, 256, 4, 4, 1024, 44, 44, (1, 1), 11, 11, 11),  # K=11, dv=256: the 512-thread cell configuration
    (1, 96, 1, 1, 3, 20, 22, (1, 1), 20, 22, 15),      # denoising shape: r=1, one head of 96, C=3
    (1, 256, 4, 4, 768, 98, 98, (1, 1), 7, 7, 7),      # C2-like cells (r=14, dv=192): tensor-core kernel, 3 value chunks
    (1, 256, 4, 4, 768, 88, 88, (1, 1), 11, 11, 11),   # K=11, dv=192: tensor-core kernel with all 512 TMEM columns
    (2, 256, 4, 4, 384, 72, 90, (1, 1), 9, 9, 9),      # the backward benchmark's K=9, dv=96 (chunks of 64 + 32), r=8x10
    (1, 256, 4, 4, 768, 56, 56, (2, 2), 4, 4, 3),      # 7-tile cells (r=28) behind replicated guidance: second Q image / S region, 3 chunks
]


@pytest.mark.parametrize("case", GRAD_CASES)
def test_fused_path_gradients_match_fp64_oracle(case):
    B, D, n_attn, n_rope, Cv, Hs, Ws, rep, h, w, K = case
    xs, feats = rnd(31, B, D, Hs, Ws), rnd(32, B, Cv, h, w)
    Ho, Wo = Hs * rep[0], Ws * rep[1]
    dout = rnd(33, B, Cv, Ho, Wo)
    x_full = xs.repeat_interleave(rep[0], 2).repeat_interleave(rep[1], 3)
    dx_full, dfeat = O.naf_forward_grads(x_full, feats, dout, n_attn, n_rope, K)
    want_dx = dx_full.reshape(B, D, Hs, rep[0], Ws, rep[1]).sum(dim=(3, 5))
    model = naf_b200.NAF(dim=D, heads_attn=n_attn, heads_rope=n_rope, kernel_size=K).eval().to(dev())
    xg = xs.to(dev()).requires_grad_(True)
    fg = feats.to(dev()).requires_grad_(True)
    out = model.upsample_from_guidance(xg, fg, rep=rep)
    with torch.no_grad():
        ref = model.upsample_from_guidance(xs.to(dev()), feats.to(dev()), rep=rep)
    assert torch.equal(out.detach(), ref)          # the differentiable forward IS the inference forward
    out.backward(dout.to(dev()))
    close(xg.grad, want_dx, (case, "dx"))
    close(fg.grad, dfeat, (case, "dfeatures"))


def test_kernel_choice_and_determinism_of_dq():
    """dq needs no atomics: bit-identical between two runs; dk / dv (atomics) agree to rounding; the cell
    and generic kernels agree with each other."""
    q, k, v, dout = rnd(1, 1, 256, 56, 56).to(dev()), rnd(2, 1, 256, 8, 8).to(dev()), rnd(3, 1, 96, 8, 8).to(dev()), \
        rnd(4, 1, 96, 56, 56).to(dev())
    a = ops.xattn_bwd(q, k, v, dout, 4, 5)
    b = ops.xattn_bwd(q, k, v, dout, 4, 5)
    g = ops.xattn_bwd(q, k, v, dout, 4, 5, algo=_lib.ALGO_GENERIC)
    assert torch.equal(a[0], b[0])
    assert ops.launch_count("xattn_bwd_cell") >= 2 and ops.launch_count("xattn_bwd_generic") >= 1
    for i in range(3):
        close(a[i], g[i].cpu(), ("cell vs generic", i))
        close(a[i], b[i].cpu(), ("run to run", i), rel=1e-5)


def test_tensor_core_backward_is_the_auto_path():
    """Integer ratios with 64-wide heads and cells of >= 64 pixels (every row of test/backward_speed.py from ratio 8
    up) run on the tensor-core cell kernel; dq is bit-reproducible, the three kernels agree."""
    q, k, v, dout = rnd(1, 2, 256, 112, 112).to(dev()) * 2, rnd(2, 2, 256, 14, 14).to(dev()), rnd(3, 2, 384, 14, 14).to(dev()), \
        rnd(4, 2, 384, 112, 112).to(dev())
    n0 = ops.launch_count("xattn_bwd_cell_tc")
    a = ops.xattn_bwd(q, k, v, dout, 4, 9)
    b = ops.xattn_bwd(q, k, v, dout, 4, 9)
    assert ops.launch_count("xattn_bwd_cell_tc") == n0 + 2
    assert torch.equal(a[0], b[0])
    s = ops.xattn_bwd(q, k, v, dout, 4, 9, algo=_lib.ALGO_CELL_SIMT)
    g = ops.xattn_bwd(q, k, v, dout, 4, 9, algo=_lib.ALGO_GENERIC)
    for i in range(3):
        close(a[i], s[i].cpu(), ("tc vs simt", i))
        close(a[i], g[i].cpu(), ("tc vs generic", i))
        close(a[i], b[i].cpu(), ("run to run", i), rel=1e-5)
    with pytest.raises(NotImplementedError):        # 49-pixel cells: not the tensor-core kernel's shape
        ops.xattn_bwd(rnd(1, 1, 256, 56, 56).to(dev()), rnd(2, 1, 256, 8, 8).to(dev()), rnd(3, 1, 96, 8, 8).to(dev()),
                      rnd(4, 1, 96, 56, 56).to(dev()), 4, 5, algo=_lib.ALGO_CELL_TC)


def test_training_step_moves_the_loss():
    """A few SGD steps through NAF (train mode: random RoPE rescale augmentation like the reference) reduce
    a regression loss: the whole differentiable path hangs together."""
    torch.manual_seed(0)
    m = naf_b200.NAF(kernel_size=7).train().to(dev())
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    img, ft = rnd(5, 2, 3, 64, 64).to(dev()), rnd(6, 2, 32, 13, 13).to(dev())
    target = rnd(7, 2, 32, 32, 32).to(dev())
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = ((m(img, ft, (32, 32)) - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert losses[-1] < losses[0]


def test_bf16_autocast_training_like_the_reference():
    """train.py:120: forward under torch.autocast(bfloat16); the gradient reaches fp32 parameters."""
    torch.manual_seed(1)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev())
    img, ft = rnd(8, 1, 3, 32, 32).to(dev()), rnd(9, 1, 16, 8, 8).to(dev()).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m(img, ft.to(torch.bfloat16), (32, 32))
    assert out.dtype == torch.bfloat16
    out.float().pow(2).mean().backward()
    assert ft.grad is not None and ft.grad.dtype == torch.float32 and torch.isfinite(ft.grad).all()
    assert all(p.grad is not None for p in m.parameters())


@pytest.mark.parametrize("cfg", [("C1", 384, 16, 224, 224, 7), ("training-like 32 <- 13", 384, 13, 56, 32, 9),
                                 ("reference backward benchmark", 384, 28, 448, 448, 9)], ids=lambda c: c[0])
def test_full_size_gradients_equal_the_reference_modules_on_the_same_gpu(cfg):
    """Forward AND backward of one full-size image through the UNMODIFIED reference modules on this GPU (oracle/_ref,
    autograd through the torch NATTEN stand-in, strict fp32) against `naf_b200.NAF` with the same weights: the output,
    d features, d image and the gradient of every parameter (what train.py:136 back-propagates)."""
    from oracle import reference_runner as R

    if not R.available():
        pytest.skip("oracle/_ref not built (python -m oracle.build_ref in the build container)")
    name, C, lo, gi, to, K = cfg
    torch.manual_seed(12)
    ref = R.load().NAF(kernel_size=K).eval().to(dev())
    ours = naf_b200.NAF(kernel_size=K).eval()
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(dev())
    image, feats, dout = rnd(31, 1, 3, gi, gi).to(dev()), rnd(32, 1, C, lo, lo).to(dev()), rnd(33, 1, C, to, to).to(dev())
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        res = []
        for model in (ref, ours):
            img, ft = image.clone().requires_grad_(True), feats.clone().requires_grad_(True)
            model.zero_grad()
            out = model(img, ft, (to, to))
            out.backward(dout)
            res.append((out.detach(), img.grad, ft.grad, {n: p.grad for n, p in model.named_parameters()}))
        (o_r, di_r, df_r, gp_r), (o_o, di_o, df_o, gp_o) = res
        assert (o_o - o_r).abs().max().item() <= 1e-4
        close(df_o, df_r.cpu(), (name, "dfeatures"), rel=1e-4)
        close(di_o, di_r.cpu(), (name, "dimage"), rel=2e-4)
        for n_, g in gp_r.items():
            assert gp_o[n_] is not None, n_
            close(gp_o[n_], g.cpu(), (name, n_), rel=3e-4)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
