"""GPU parity tests: the CUDA path (through the C ABI) against the reference-generated golden
fixtures, against the CPU oracle on seeded inputs, and -- at the full BASELINE.json sizes --
through size-independent properties.

Tolerances: fp32 path; BASELINE.json's bar is <= 1e-3 max-abs.  The SIMT/generic kernels are
held to a much tighter 2e-5 (they are plain fp32 with a different summation order); index
tables must be bit-exact.
"""
import numpy as np
import pytest
import torch

import golden_util as G
import naf_b200
from naf_b200 import _lib, ops, taps
from oracle import naf_oracle as O

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-5      # fp32 kernels vs CPU fp32 oracle
TOL_BAR = 1e-3       # BASELINE.json north_star tolerance (tensor-core path must meet this)


def dev():
    return torch.device("cuda", 0)


def tol_for(algo):
    return TOL_BAR if algo in (_lib.ALGO_CELL_TCWS, _lib.ALGO_CELL_TMA, _lib.ALGO_UNION_TC) else TOL_FP32


def available_algos(q_shape, v_shape, heads, K):
    """Every kernel able to run this problem (each is tested, not just AUTO's pick)."""
    import ctypes as C
    algos = [_lib.ALGO_GENERIC]
    Ho, Wo, h, w = q_shape[2], q_shape[3], v_shape[2], v_shape[3]
    integer = Ho % h == 0 and Wo % w == 0
    for algo in (_lib.ALGO_CELL_SIMT, _lib.ALGO_CELL_TCWS, _lib.ALGO_CELL_TMA, _lib.ALGO_UNION_TC):
        if not integer and algo != _lib.ALGO_UNION_TC:
            continue
        p = _lib.XAttnParams()
        d = 1 << 20
        p.q = p.k = p.v = p.out = d
        if not integer:
            p.row_tap = p.col_tap = d
        p.workspace, p.workspace_bytes = d, 1 << 62
        p.B, p.D, p.C, p.heads = q_shape[0], q_shape[1], v_shape[1], heads
        p.Ho, p.Wo, p.h, p.w, p.K = Ho, Wo, h, w, K
        p.scale = 1.0
        p.q_stride_b, p.q_stride_y, p.q_stride_x = Ho * Wo * q_shape[1], Wo * q_shape[1], q_shape[1]
        p.algo = algo
        if _lib.load().naf_xattn_select_algo(C.byref(p)) == algo:
            algos.append(algo)
    return algos


# ------------------------------------------------------------------ golden fixtures (reference)
@pytest.mark.parametrize("name", G.names("xattn_"))
def test_cross_attention_matches_reference_golden(name):
    c = G.attention_case(name)
    D = c["q"].shape[1]
    mod = naf_b200.CrossAttention(dim=D, num_heads=c["heads"], kernel_size=(c["K"], c["K"]))
    q, k, v = (c[n].to(dev()) for n in "qkv")
    for algo in available_algos(q.shape, v.shape, c["heads"], c["K"]):
        mod.algo = algo
        out = mod(q, k, v, None)
        assert out.shape == c["out"].shape
        err = (out.cpu() - c["out"]).abs().max().item()
        assert err <= tol_for(algo), (name, _lib.ALGO_NAMES[algo], err)
        assert mod.dilation == c["dilation"]
    if c["scores"] is not None:
        # return_weights=True: the generic kernel, and the TMA cell kernel where it takes the shape (scores
        # come straight from the tensor-core logits there: fp32-class through the 3-pass split, 1e-4 bar)
        algos = [_lib.ALGO_GENERIC]
        if _lib.ALGO_CELL_TMA in available_algos(q.shape, v.shape, c["heads"], c["K"]):
            algos.append(_lib.ALGO_CELL_TMA)
        for algo in algos:
            mod.algo = algo
            out, scores = mod(q, k, v, None, return_weights=True)
            assert scores.shape == c["scores"].shape
            tc = algo == _lib.ALGO_CELL_TMA
            assert (scores.cpu() - c["scores"]).abs().max().item() <= (1e-4 if tc else 2e-5), _lib.ALGO_NAMES[algo]
            assert (out.cpu() - c["out"]).abs().max().item() <= tol_for(algo)


@pytest.mark.parametrize("name", G.names("rope_"))
def test_rope_matches_reference_golden(name):
    c = G.rope_case(name)
    D = c["x"].shape[1]
    rope = naf_b200.RoPE(D, num_heads=c["heads"], base=100.0, rescale_coords=2.0).eval().to(dev())
    out = rope(c["x"].to(dev()))
    # cos/sin come from the CUDA libm instead of the CPU one: a few ulp
    assert (out.cpu() - c["out"]).abs().max().item() <= 5e-6
    flat = rope(c["x"].to(dev()), layout="flatten")
    assert flat.shape == (c["x"].shape[0], c["x"].shape[2], c["x"].shape[3], c["heads"], D // c["heads"])


@pytest.mark.parametrize("name", G.names("naf_"))
def test_naf_module_matches_reference_golden(name):
    """Whole NAF.forward (cuDNN encoder in strict fp32 + our kernels) vs the reference."""
    c = G.module_case(name)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = naf_b200.NAF(dim=128, kernel_size=7).eval()
        m.load_state_dict(G.module_state())
        m = m.to(dev())
        out = m(c["image"].to(dev()), c["features"].to(dev()), c["output_size"])
        assert out.shape == c["out"].shape
        assert (out.cpu() - c["out"]).abs().max().item() <= 1e-4, name
        q = m.image_encoder(c["image"].to(dev()), c["output_size"])
        assert (q.cpu() - c["queries"]).abs().max().item() <= 1e-4
        k = m.key_encoder(q, c["features"].to(dev()))
        assert (k.cpu() - O.key_pool(c["queries"], *c["features"].shape[-2:])).abs().max().item() <= 1e-4
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------ whole images against the reference itself
@pytest.mark.parametrize("cfg", [("C1", 384, 16, 224, 224, 7), ("C2", 768, 32, 448, 896, 7), ("C3", 1024, 37, 518, 1036, 11),
                                 ("C4", 768, 24, 336, 1344, 7), ("C5", 768, 32, 512, 2048, 7),
                                 ("C2 image at target size", 768, 32, 896, 896, 7)], ids=lambda c: c[0])
def test_full_size_image_equals_the_reference_modules_on_the_same_gpu(cfg):
    """One FULL-SIZE image of a BASELINE config through the UNMODIFIED reference modules (oracle/_ref: src/model/naf.py
    NAF.forward, NATTEN's two functionals replaced by the pure-torch stand-in) run on this GPU in strict fp32, against
    `naf_b200.NAF` with the same weights: every output element of the whole image, both precision classes of the
    encoder.  (The CPU oracle is too slow at these sizes; this is the reference's own code on the same inputs.)"""
    from oracle import reference_runner as R

    if not R.available():
        pytest.skip("oracle/_ref not built (python -m oracle.build_ref in the build container)")
    name, C, lo, gi, to, K = cfg
    torch.manual_seed(11)
    ref = R.load().NAF(kernel_size=K).eval().to(dev())
    ours = naf_b200.NAF(kernel_size=K).eval()
    ours.load_state_dict(ref.state_dict())
    ours = ours.to(dev())
    image, feats = rnd(21, 1, 3, gi, gi).to(dev()), rnd(22, 1, C, lo, lo).to(dev())
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.no_grad():
            want = ref(image, feats, (to, to))
            got = ours(image, feats, (to, to))          # strict class: split-fp16 three-pass encoder
        assert got.shape == want.shape == (1, C, to, to)
        err = (got - want).abs().max().item()
        assert err <= 1e-4, (name, "strict", err)
        torch.backends.cudnn.allow_tf32 = True            # PyTorch's default = what the bench runs
        with torch.no_grad():
            got_tf32 = ours(image, feats, (to, to))
        err_tf32 = (got_tf32 - want).abs().max().item()
        assert err_tf32 <= 1e-3, (name, "tf32 class", err_tf32)
        print(f"{name}: max|ours - reference| strict {err:.2e}, tf32 class {err_tf32:.2e}")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------ rectangular windows
@pytest.mark.parametrize("name", G.names("rect_xattn_"))
def test_rectangular_windows_match_reference_golden(name):
    """kernel_size=(kh, kw) with kh != kw, as NATTEN takes it from the reference (src/layers/attentions.py:20,24):
    forward, scores (tap order t_h*kw + t_w) and gradients against fixtures from the unmodified reference.  AUTO
    routes these to the generic kernels (fp32: 2e-5); asking a cell kernel for one fails loudly."""
    c = G.rect_attention_case(name)
    D = c["q"].shape[1]
    mod = naf_b200.CrossAttention(dim=D, num_heads=c["heads"], kernel_size=c["K"])
    q, k, v, dout = (c[n].to(dev()) for n in ("q", "k", "v", "dout"))
    assert ops.select_algo(q.shape, v.shape, c["heads"], c["K"], rope_on_the_fly=False) == "generic"
    with torch.no_grad():
        out = mod(q, k, v, None)
        assert mod.dilation == c["dilation"]
        assert (out.cpu() - c["out"]).abs().max().item() <= 2e-5
        out, scores = mod(q, k, v, None, return_weights=True)
        assert scores.shape == c["scores"].shape
        assert (scores.cpu() - c["scores"]).abs().max().item() <= 2e-5
        assert (out.cpu() - c["out"]).abs().max().item() <= 2e-5
        for algo in (_lib.ALGO_CELL_SIMT, _lib.ALGO_CELL_TCWS, _lib.ALGO_CELL_TMA, _lib.ALGO_UNION_TC):
            with pytest.raises(NotImplementedError):
                ops.xattn(q, k, v, c["heads"], c["K"], algo=algo)
    # gradients: the operator-level entry point and autograd through the module
    dq, dk, dv = ops.xattn_bwd(q, k, v, dout, c["heads"], c["K"])
    qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
    mod(qg, kg, vg, None).backward(dout)
    for got, got2, want, what in ((dq, qg.grad, c["dq"], "dq"), (dk, kg.grad, c["dk"], "dk"), (dv, vg.grad, c["dv"], "dv")):
        scale = max(want.abs().max().item(), 1e-6)
        assert (got.cpu() - want).abs().max().item() <= 3e-5 * scale, (name, what)
        assert (got2.cpu() - want).abs().max().item() <= 3e-5 * scale, (name, what, "autograd")
    with pytest.raises(NotImplementedError):
        ops.xattn_bwd(q, k, v, dout, c["heads"], c["K"], algo=_lib.ALGO_CELL_SIMT)


def test_rectangular_window_with_rope_and_replication_matches_oracle():
    """The fused form NAF.forward uses (un-rotated guidance + rope tables, replicated source map) with a 5 x 9 window."""
    torch.manual_seed(3)
    B, D, C, heads, h, w, r, K = 1, 256, 32, 4, 9, 10, 4, (5, 9)
    Ho, Wo = h * r, w * r
    m = naf_b200.NAF(kernel_size=7).eval().to(dev())
    m.upsampler.kernel_size = K
    xs = torch.randn(B, D, Ho // 2, Wo // 2)
    feats = torch.randn(B, C, h, w)
    x_full = xs.repeat_interleave(2, 2).repeat_interleave(2, 3)
    want = O.naf_forward(x_full, feats, heads, 4, K)
    tables = m.image_encoder.rope.axis_tables(Ho, Wo)
    with torch.no_grad():
        kk, _ = ops.rope_kpool(xs.to(dev()), tables, 4, pooled_hw=(h, w), rep=(2, 2))
        out = m.upsampler(xs.to(dev()), kk, feats.to(dev()), rope_tables=tables, rep=(2, 2))
    assert (out.cpu() - want).abs().max().item() <= 2e-5


# ------------------------------------------------------------------ bit-exact neighbourhood indices
@pytest.mark.parametrize("shape", [(36, 36, 9, 9, 7), (224, 224, 16, 16, 7), (1036, 1036, 37, 37, 11),
                                   (24, 40, 8, 10, 5), (32, 32, 13, 13, 9), (30, 45, 7, 11, 3),
                                   (20, 22, 20, 22, 15), (896, 896, 32, 32, 7)])
def test_device_tap_indices_bit_exact(shape):
    Ho, Wo, h, w, K = shape
    rt, ct = O.tap_tables(Ho, Wo, h, w, K)
    want = (rt.astype(np.int64)[:, None, :, None] * w + ct.astype(np.int64)[None, :, None, :]).reshape(Ho, Wo, K * K)
    got = ops.dump_taps(Ho, Wo, h, w, K, dev()).cpu().numpy()
    assert (got == want).all()
    got_tab = ops.dump_taps(Ho, Wo, h, w, K, dev(), force_tables=True).cpu().numpy()
    assert (got_tab == want).all()


# ------------------------------------------------------------------ seeded oracle comparisons
def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


ORACLE_CASES = [
    # B, D, n, C, Ho, Wo, h, w, K, gain
    (1, 256, 4, 384, 112, 112, 8, 8, 7, 1.0),     # C1-like (r=14)
    (1, 256, 4, 768, 196, 196, 7, 7, 7, 3.0),     # C2-like (r=28, dv=192), peaky
    (1, 256, 4, 1024, 154, 154, 11, 11, 11, 1.0),  # C3-like (K=11, dv=256, r=14)
    (2, 256, 4, 128, 64, 96, 8, 12, 5, 2.0),      # dv=32, non-square, batch 2
    (1, 128, 4, 96, 63, 63, 9, 9, 9, 1.0),        # dq=32, dv=24, odd ratio 7
    (1, 64, 4, 200, 48, 48, 8, 8, 3, 1.0),        # dq=16, dv=50 (generic only)
    (1, 256, 4, 60, 45, 45, 9, 9, 5, 1.0),        # r=5: cells of 25 pixels (small-cell path)
    (1, 256, 2, 6, 40, 40, 8, 8, 7, 8.0),         # dq=128 (generic), very peaky
    # tap-table / ratio-1 / tiny-cell shapes: the union-window tensor-core kernel (and the generic one)
    (2, 256, 4, 384, 32, 32, 13, 13, 9, 2.0),     # the reference's training shape 32 <- 13 (duplicated taps)
    (1, 256, 4, 768, 32, 32, 19, 19, 9, 1.0),     # 32 <- 19: ratio 1.68, dilation 1
    (1, 256, 4, 1024, 37, 41, 16, 18, 9, 3.0),    # dv = 256, non-square, partial tiles
    (2, 96, 1, 3, 40, 52, 40, 52, 15, 4.0),       # denoising: ratio 1, C = 3, one head, K = 15 (6 chunks)
    (1, 256, 1, 3, 48, 48, 48, 48, 11, 2.0),      # denoising dim 256: dq = 256
    (1, 128, 1, 4, 33, 35, 33, 35, 7, 1.0),       # heads 1-4 variant, odd sizes
    (1, 256, 4, 384, 56, 56, 28, 28, 9, 2.0),     # integer ratio 2: 4-pixel cells
    (1, 64, 2, 10, 30, 45, 7, 11, 3, 4.0),        # ratio 4.3 / 4.1, dv = 5
    (1, 256, 4, 200, 45, 33, 9, 11, 5, 1.0),      # dv = 50: padded value head, scalar loads
]


@pytest.mark.parametrize("case", ORACLE_CASES)
def test_kernels_match_oracle(case):
    B, D, n, Cv, Ho, Wo, h, w, K, gain = case
    q, k, v = rnd(1, B, D, Ho, Wo) * gain, rnd(2, B, D, h, w), rnd(3, B, Cv, h, w)
    want = O.cross_attention(q, k, v, n, K)
    for algo in available_algos(q.shape, v.shape, n, K):
        got = ops.xattn(q.to(dev()), k.to(dev()), v.to(dev()), n, K, algo=algo)
        err = (got.cpu() - want).abs().max().item()
        assert err <= tol_for(algo), (case, _lib.ALGO_NAMES[algo], err)


def test_return_weights_from_the_tensor_core_kernel():
    """return_weights=True on a C1-shaped problem is served by the TMA cell kernel (not the warp-per-pixel
    generic one) and returns the oracle's scaled pre-softmax scores, NATTEN tap order."""
    B, D, n, Cv, Ho, h, K = 2, 256, 4, 384, 112, 8, 7
    q, k, v = rnd(1, B, D, Ho, Ho) * 2.0, rnd(2, B, D, h, h), rnd(3, B, Cv, h, h)
    want, want_s = O.cross_attention(q, k, v, n, K, return_weights=True)
    assert ops.select_algo(q.shape, v.shape, n, K, rope_on_the_fly=False, return_scores=True) == "cell_tma"
    n0 = ops.launch_count("xattn_cell_tma")
    got, got_s = ops.xattn(q.to(dev()), k.to(dev()), v.to(dev()), n, K, return_scores=True)
    assert ops.launch_count("xattn_cell_tma") == n0 + 1
    assert got_s.shape == want_s.shape == (B, n, Ho, Ho, K * K)
    assert (got_s.cpu() - want_s).abs().max().item() <= 1e-4
    assert (got.cpu() - want).abs().max().item() <= TOL_BAR


@pytest.mark.parametrize("case", [(1, 256, 4, 4, 384, 112, 112, 8, 8, 7), (2, 128, 4, 4, 64, 56, 84, 8, 12, 5),
                                  (1, 256, 4, 4, 16, 32, 32, 13, 13, 9), (1, 256, 4, 1, 32, 60, 60, 10, 10, 3),
                                  (1, 96, 1, 1, 3, 20, 22, 20, 22, 15)])
def test_fused_rope_path_matches_oracle(case):
    """x (un-rotated) -> rope+kpool kernel -> attention with on-the-fly RoPE (or materialised q
    when rope heads != attention heads) == oracle naf_forward."""
    B, D, n_attn, n_rope, Cv, Ho, Wo, h, w, K = case
    x, feats = rnd(4, B, D, Ho, Wo), rnd(5, B, Cv, h, w)
    want = O.naf_forward(x, feats, n_attn, n_rope, K)
    model = naf_b200.NAF(dim=D, heads_attn=n_attn, heads_rope=n_rope, kernel_size=K).eval().to(dev())
    for layout in ("nchw", "channels_last"):
        xd = x.to(dev())
        if layout == "channels_last":
            xd = xd.contiguous(memory_format=torch.channels_last)
        got = model.upsample_from_guidance(xd, feats.to(dev()))
        err = (got.cpu() - want).abs().max().item()
        assert err <= 5e-5, (case, layout, err)


@pytest.mark.parametrize("rep", [(2, 2), (1, 3), (4, 2)])
def test_replicated_guidance_source(rep):
    """Kernels reading the encoder-resolution map through `rep` == oracle on the replicated map
    (what the reference's adaptive_avg_pool2d produces when the target is a multiple)."""
    B, D, Cv, Hs, Ws, K = 2, 256, 128, 42, 56, 3
    Ho, Wo = Hs * rep[0], Ws * rep[1]
    h, w = Ho // 14, Wo // 14
    xs, feats = rnd(7, B, D, Hs, Ws), rnd(8, B, Cv, h, w)
    x_full = torch.nn.functional.adaptive_avg_pool2d(xs, (Ho, Wo))
    assert torch.equal(x_full, xs.repeat_interleave(rep[0], 2).repeat_interleave(rep[1], 3))
    want = O.naf_forward(x_full, feats, 4, 4, K)
    model = naf_b200.NAF(kernel_size=K).eval().to(dev())
    for algo in (_lib.ALGO_GENERIC, _lib.ALGO_AUTO):
        model.upsampler.algo = algo
        got = model.upsample_from_guidance(xs.to(dev()), feats.to(dev()), rep=rep)
        assert got.shape == want.shape
        assert (got.cpu() - want).abs().max().item() <= 5e-5


def test_strided_and_unaligned_inputs():
    """Views with odd strides / offsets go through the packing kernel and still match."""
    B, D, n, Cv, Ho, h, K = 1, 64, 4, 20, 28, 7, 3
    qbig, kbig, vbig = rnd(1, B, D, Ho + 3, Ho + 1), rnd(2, B, D + 1, h, h), rnd(3, B, Cv, h, h + 2)
    q, k, v = qbig[:, :, 1:Ho + 1, 1:], kbig[:, 1:], vbig[..., 1:h + 1]
    want = O.cross_attention(q.contiguous(), k.contiguous(), v.contiguous(), n, K)
    got = ops.xattn(qbig.to(dev())[:, :, 1:Ho + 1, 1:], kbig.to(dev())[:, 1:], vbig.to(dev())[..., 1:h + 1], n, K)
    assert (got.cpu() - want).abs().max().item() <= TOL_FP32
    # token-major ViT features: (B, hw, C) -> permuted NCHW view
    tok = rnd(6, B, h * h, Cv)
    v2 = tok.permute(0, 2, 1).reshape(B, Cv, h, h)
    want2 = O.cross_attention(q.contiguous(), k.contiguous(), v2.contiguous(), n, K)
    got2 = ops.xattn(q.contiguous().to(dev()), k.contiguous().to(dev()),
                     tok.to(dev()).permute(0, 2, 1).reshape(B, Cv, h, h), n, K)
    assert (got2.cpu() - want2).abs().max().item() <= TOL_FP32


def test_output_is_pixel_major_view_like_the_reference():
    q, k, v = rnd(1, 1, 64, 16, 16).to(dev()), rnd(2, 1, 64, 4, 4).to(dev()), rnd(3, 1, 8, 4, 4).to(dev())
    out = ops.xattn(q, k, v, 4, 3)
    assert out.shape == (1, 8, 16, 16) and out.stride(1) == 1 and not out.is_contiguous()
    assert out.permute(0, 2, 3, 1).is_contiguous()


def test_errors_on_device():
    q, k, v = torch.zeros(1, 64, 16, 16, device=dev()), torch.zeros(1, 64, 4, 4, device=dev()), torch.zeros(1, 8, 4, 4, device=dev())
    with pytest.raises(ValueError):
        ops.xattn(q, k, v, 4, 5)      # K*dilation = 20 > 16
    with pytest.raises(ValueError):
        ops.xattn(q, k, v, 4, 2)      # even kernel
    with pytest.raises(ValueError):
        ops.xattn(q, k, torch.zeros(1, 6, 4, 4, device=dev()), 4, 3)  # C % heads
    with pytest.raises(RuntimeError, match="forward pass only"):
        ops.xattn(q.requires_grad_(), k, v, 4, 3)
    with pytest.raises(ValueError):
        naf_b200.CrossAttention(64, 4, (3, 5))(q.detach(), k, v)      # rectangular: kw*dilation = 20 > 16
    with pytest.raises(ValueError):
        naf_b200.CrossAttention(64, 4, (3, 4))(q.detach(), k, v)      # even width
    assert naf_b200.CrossAttention(64, 4, (1, 3))(q.detach(), k, v).shape == (1, 8, 16, 16)


# ------------------------------------------------------------------ full-size property tests
FULL = [  # B(1 image of each BASELINE config), C, h, Ho, K
    ("C1", 384, 16, 224, 7), ("C2", 768, 32, 896, 7), ("C3", 1024, 37, 1036, 11),
    ("C4", 768, 24, 1344, 7), ("C5", 768, 32, 2048, 7),
]


@pytest.mark.parametrize("cfg", FULL, ids=[c[0] for c in FULL])
def test_full_size_properties(cfg):
    """At the real sizes the oracle is too slow, so check properties that pin the result:
    (1) constant V  -> output constant; (2) linearity in V; (3) zero queries -> the output of a
    cell is the mean of its clamped window (computed with an integral image on the GPU);
    (4) spot-check random pixels against the oracle formula evaluated for those pixels only."""
    _, Cv, h, Ho, K = cfg
    n, D = 4, 256
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(1, D, Ho, Ho, generator=g).to(dev()).contiguous(memory_format=torch.channels_last)
    v1 = torch.randn(1, Cv, h, h, generator=g).to(dev())
    v2 = torch.randn(1, Cv, h, h, generator=g).to(dev())
    model = naf_b200.NAF(kernel_size=K).eval().to(dev())
    rope = model.image_encoder.rope
    tables = rope.axis_tables(Ho, Ho)
    k, _ = ops.rope_kpool(x, tables, 4, pooled_hw=(h, h))

    def run(v, q=x, rope_tables=tables):
        return ops.xattn(q, k, v, n, K, rope_tables=rope_tables)

    # (1) constant values
    out = run(torch.full_like(v1, 1.5))
    assert (out - 1.5).abs().max().item() <= 1e-5
    del out
    # (2) linearity in V
    o1, o2 = run(v1), run(v2)
    o12 = run(2.0 * v1 - 0.5 * v2)
    assert (o12 - (2.0 * o1 - 0.5 * o2)).abs().max().item() <= 1e-4
    del o2, o12
    # (4) spot check against the oracle formula at random pixels
    r = Ho // h
    rt, ct = O.tap_tables(Ho, Ho, h, h, K)
    qrot = O.rope_rotate  # noqa: F841  (documentational: q below is rotated on the GPU)
    _, qfull = ops.rope_kpool(x[:, :, :r * 2], (tables[0][:r * 2], tables[1][:r * 2], tables[2], tables[3]), 4,
                              pooled_hw=None, want_q=True)
    rs = np.random.RandomState(5)
    kc, vc = k.cpu(), v1.cpu()
    for _ in range(24):
        y, xx = int(rs.randint(0, 2 * r)), int(rs.randint(0, Ho))
        qv = qfull[0, :, y, xx].cpu().view(n, 64)
        rows = torch.from_numpy(rt[y].astype(np.int64))
        cols = torch.from_numpy(ct[xx].astype(np.int64))
        kw = kc[0][:, rows][:, :, cols].reshape(n, 64, K * K)
        vw = vc[0][:, rows][:, :, cols].reshape(n, Cv // n, K * K)
        p = torch.softmax(torch.einsum("nd,ndt->nt", qv, kw) * 0.125, dim=-1)
        want = torch.einsum("nt,nct->nc", p, vw).reshape(-1)
        assert (o1[0, :, y, xx].cpu() - want).abs().max().item() <= 1e-4
    del o1
    # (3) zero queries: uniform softmax -> window mean
    out0 = run(v1, q=torch.zeros_like(x), rope_tables=None)
    origin = torch.clamp(torch.arange(h, device=dev()) - K // 2, 0, h - K)
    win = v1.unfold(2, K, 1).unfold(3, K, 1).mean(dim=(-1, -2))  # (1,C,h-K+1,h-K+1)
    want0 = win[:, :, origin][:, :, :, origin]                   # per cell
    got0 = out0.reshape(1, Cv, h, r, h, r)
    assert (got0 - want0[:, :, :, None, :, None]).abs().max().item() <= 1e-5


# ------------------------------------------------------------------ full-size, whole image + whole batch
# name, C, low-res side, guidance-source side, target side, K, batch (second half = copies of the first)
FULL_BATCH = [
    ("C1", 384, 16, 224, 224, 7, 4),
    ("C2", 768, 32, 448, 896, 7, 8),       # the bench launch: 4.93 G output elements (> 2^32)
    ("C3", 1024, 37, 518, 1036, 11, 4),    # one GPU's shard of the 8-GPU config
    ("C4", 768, 24, 336, 1344, 7, 6),      # 8.3 G output elements
    ("C5", 768, 32, 512, 2048, 7, 4),      # 12.9 G output elements
]


def _spot_pixels(Ho, r, rs, n_random=16):
    """Target pixels spread over the WHOLE image: the four corners, the last two cell rows / columns
    (bottom / right window clamping), the first two, and random interior pixels."""
    edge = [0, 1, r - 1, r, 2 * r - 1, Ho - 2 * r, Ho - r - 1, Ho - r, Ho - 2, Ho - 1]
    pts = [(y, x) for y in (0, Ho - 1) for x in (0, Ho - 1)]
    pts += [(int(rs.choice(edge)), int(rs.randint(0, Ho))) for _ in range(8)]
    pts += [(int(rs.randint(0, Ho)), int(rs.choice(edge))) for _ in range(8)]
    pts += [(int(rs.choice(edge)), int(rs.choice(edge))) for _ in range(6)]
    pts += [(int(rs.randint(0, Ho)), int(rs.randint(0, Ho))) for _ in range(n_random)]
    return pts


@pytest.mark.parametrize("cfg", FULL_BATCH, ids=[c[0] for c in FULL_BATCH])
def test_full_batch_launch_whole_image(cfg):
    """One launch at the full BASELINE size and batch, exactly as `NAF.forward` issues it (guidance map
    at the encoder resolution read through replication factors, RoPE fused):
    (1) 64-bit indexing: images B/2.. are copies of images 0..B/2-1, so their outputs must be
        bit-identical although they live beyond 2^31 (C2, C4, C5: beyond 2^32) elements;
    (2) batch-position independence: image 0 of the batch == the same image run alone, bit for bit;
    (3) keys: the pooled keys of cells in every corner == block mean of the oracle's rotation;
    (4) spot checks of the oracle formula over the whole image, last two cell rows / columns included,
        for the first and the last image of the batch."""
    name, Cv, h, Hs, Ho, K, B = cfg
    n, D, rep = 4, 256, Ho // Hs
    r = Ho // h
    g = torch.Generator(device="cpu").manual_seed(21)
    half = B // 2
    xs_h = torch.randn(half, Hs, Hs, D, generator=g)
    v_h = torch.randn(half, h, h, Cv, generator=g)
    xs = torch.cat([xs_h, xs_h]).to(dev()).permute(0, 3, 1, 2)       # pixel-major storage, NCHW-shaped
    v = torch.cat([v_h, v_h]).to(dev()).permute(0, 3, 1, 2)
    model = naf_b200.NAF(kernel_size=K).eval().to(dev())
    out = model.upsample_from_guidance(xs, v, rep=(rep, rep))
    assert out.shape == (B, Cv, Ho, Ho)
    assert out.numel() > 2 ** 31 or name in ("C1",)
    # (1)
    for i in range(half):
        assert torch.equal(out[i + half], out[i]), (name, i)
    # (2)
    alone = model.upsample_from_guidance(xs[:1], v[:1], rep=(rep, rep))
    assert torch.equal(alone[0], out[0])
    del alone
    # (3) + (4) against the oracle formulas
    rope = model.image_encoder.rope
    per = O.rope_periods(64)
    assert torch.equal(per, rope.periods.cpu())
    k = ops.rope_kpool(xs, rope.mean_axis_tables(Ho, Ho, rep, rep), 4, pooled_hw=(h, h))[0].cpu()   # (B,D,h,h)
    rs = np.random.RandomState(9)
    for (ci, cj) in [(0, 0), (0, h - 1), (h - 1, 0), (h - 1, h - 1), (h // 2, h // 3)]:
        yy, xx = np.meshgrid(np.arange(ci * r, ci * r + r), np.arange(cj * r, cj * r + r), indexing="ij")
        yy, xx = yy.reshape(-1), xx.reshape(-1)
        src = xs_h[0][torch.from_numpy(yy // rep), torch.from_numpy(xx // rep)].double()
        want_k = O.rope_rotate_pixels(src, yy, xx, Ho, Ho, 4, per.double()).mean(0)
        assert (k[0, :, ci, cj].double() - want_k).abs().max().item() <= 2e-5, (name, ci, cj)
    rt, ct = O.tap_tables(Ho, Ho, h, h, K)
    for b in (0, B - 1):
        pts = _spot_pixels(Ho, r, rs)
        ys, xx = [p[0] for p in pts], [p[1] for p in pts]
        src = xs_h[b % half][torch.tensor(ys) // rep, torch.tensor(xx) // rep]
        qrot = O.rope_rotate_pixels(src, ys, xx, Ho, Ho, 4, per)          # (N, D) fp32, oracle formula
        got = out[b][:, torch.tensor(ys, device=dev()), torch.tensor(xx, device=dev())].cpu()   # (C, N)
        kc, vc = k[b], v_h[b % half].permute(2, 0, 1)
        for i, (y, x) in enumerate(pts):
            rows = torch.from_numpy(rt[y].astype(np.int64))
            cols = torch.from_numpy(ct[x].astype(np.int64))
            kw = kc[:, rows][:, :, cols].reshape(n, 64, K * K)
            vw = vc[:, rows][:, :, cols].reshape(n, Cv // n, K * K)
            p = torch.softmax(torch.einsum("nd,ndt->nt", qrot[i].view(n, 64), kw) * 0.125, dim=-1)
            want = torch.einsum("nt,nct->nc", p, vw).reshape(-1)
            assert (got[:, i] - want).abs().max().item() <= 1e-4, (name, b, y, x)


# ------------------------------------------------------------------ bf16 output
@pytest.mark.parametrize("case", [(1, 256, 4, 768, 196, 196, 7, 7, 7), (2, 256, 4, 128, 64, 96, 8, 12, 5),
                                  (1, 256, 4, 1024, 154, 154, 11, 11, 11), (1, 64, 4, 16, 32, 32, 13, 13, 9)])
def test_bf16_output_is_the_rounded_fp32_output(case):
    """out_dtype=bfloat16 changes only the final store: bit-identical to rounding the fp32 result."""
    B, D, n, C, Ho, Wo, h, w, K = case
    q, k, v = rnd(1, B, D, Ho, Wo).to(dev()), rnd(2, B, D, h, w).to(dev()), rnd(3, B, C, h, w).to(dev())
    for algo in available_algos(q.shape, v.shape, n, K):
        if algo == _lib.ALGO_CELL_SIMT:
            with pytest.raises(NotImplementedError):
                ops.xattn(q, k, v, n, K, algo=algo, out_dtype=torch.bfloat16)
            continue
        o32 = ops.xattn(q, k, v, n, K, algo=algo)
        try:
            o16 = ops.xattn(q, k, v, n, K, algo=algo, out_dtype=torch.bfloat16)
        except NotImplementedError:
            # the TMA kernel stores bf16 in 64-channel boxes: value heads of 32 / 96 channels stay on the
            # other kernels (AUTO falls through to them)
            assert algo == _lib.ALGO_CELL_TMA and (C // n) % 64 != 0
            continue
        assert o16.dtype == torch.bfloat16 and o16.shape == o32.shape and o16.stride() == o32.stride()
        if not torch.equal(o16, o32.to(torch.bfloat16)):
            bad = o16 != o32.to(torch.bfloat16)
            o32b = ops.xattn(q, k, v, n, K, algo=algo)
            o16b = ops.xattn(q, k, v, n, K, algo=algo, out_dtype=torch.bfloat16)
            raise AssertionError(
                f"{_lib.ALGO_NAMES[algo]}: {bad.float().mean().item():.4f} of the bf16 elements differ; per-channel "
                f"{[round(x, 2) for x in bad.float().mean(dim=(0, 2, 3)).tolist()[::16]]}; per-row "
                f"{[round(x, 2) for x in bad.float().mean(dim=(0, 1, 3)).tolist()[::8]]}; rerun: fp32 repeatable "
                f"{torch.equal(o32, o32b)}, bf16 now right {torch.equal(o16b, o32b.to(torch.bfloat16))}, "
                f"bf16 repeatable {torch.equal(o16, o16b)}")


def test_module_follows_bf16_autocast_like_the_reference():
    m = naf_b200.NAF(kernel_size=7).eval().to(dev())
    img, ft = rnd(4, 1, 3, 64, 64).to(dev()), rnd(5, 1, 32, 8, 8).to(dev())
    o32 = m(img, ft, (64, 64))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o16 = m(img, ft.to(torch.bfloat16), (64, 64))
    assert o32.dtype == torch.float32 and o16.dtype == torch.bfloat16
    o32b = m(img, ft.to(torch.bfloat16).float(), (64, 64))   # same rounded features, fp32 out
    # (a different kernel may serve the bf16 request: equal up to one bf16 rounding step)
    assert ((o16.float() - o32b).abs() <= 2.0 ** -7 * o32b.abs() + 1e-6).all()
    assert m(img, ft, (64, 64), out_dtype=torch.bfloat16).dtype == torch.bfloat16


# ------------------------------------------------------------------ bf16 inputs
@pytest.mark.parametrize("case", [(2, 256, 4, 768, 112, 112, 8, 8, 7), (1, 256, 4, 1024, 154, 154, 11, 11, 11),
                                  (1, 256, 4, 384, 126, 98, 9, 7, 5)])
@pytest.mark.parametrize("layout", ["nchw", "channels_last"])
def test_bf16_inputs_are_read_natively(case, layout):
    """bf16 q / k / v (what the reference's NATTEN calls see under torch.autocast(bfloat16), train.py:120) go
    to the TMA kernel as they are -- no widened fp32 copy -- and give bit for bit the result of the fp32
    path on the same (bf16-representable) values; mixed dtypes too (fp32 guidance + bf16 features is what
    NAF.forward itself produces under autocast)."""
    B, D, n, Cv, Ho, Wo, h, w, K = case
    q, k, v = (t.to(torch.bfloat16).to(dev()) for t in (rnd(1, B, D, Ho, Wo), rnd(2, B, D, h, w), rnd(3, B, Cv, h, w)))
    if layout == "channels_last":
        q, k, v = (t.contiguous(memory_format=torch.channels_last) for t in (q, k, v))
    want = ops.xattn(q.float(), k.float(), v.float(), n, K, algo=_lib.ALGO_CELL_TMA)
    n0 = ops.launch_count("pack_nhwc")
    got = ops.xattn(q, k, v, n, K)
    if layout == "channels_last":
        assert ops.launch_count("pack_nhwc") == n0          # nothing was repacked or widened by our kernels
    assert got.dtype == torch.float32 and torch.equal(got, want)
    mixed = ops.xattn(q.float(), k.float(), v, n, K)
    assert torch.equal(mixed, want)
    got16 = ops.xattn(q, k, v, n, K, out_dtype=torch.bfloat16)
    if (Cv // n) % 64 == 0:
        assert torch.equal(got16, want.to(torch.bfloat16))
    # the other kernels take fp32 only: a forced kernel still works (the Python layer widens for it)
    gen = ops.xattn(q, k, v, n, K, algo=_lib.ALGO_GENERIC)
    assert (gen - want).abs().max().item() <= TOL_BAR


# ------------------------------------------------------------------ union-window tensor-core kernel
def test_union_kernel_is_the_auto_path_for_tap_tables():
    """Non-integer ratios (the reference's training shape 32 <- 13, utils/training.py:28-50) and ratio 1
    (denoising.py:209-213) run on the tensor-core union kernel under AUTO -- not on the warp-per-pixel kernel --
    with bf16 stores and on-the-fly RoPE, and agree with the fp32 generic kernel to the 1e-3 bar (measured 1e-5)."""
    for (B, D, n, Cv, Ho, Wo, h, w, K) in [(2, 256, 4, 384, 32, 32, 13, 13, 9), (1, 96, 1, 3, 40, 36, 40, 36, 15)]:
        q, k, v = rnd(1, B, D, Ho, Wo).to(dev()) * 2.0, rnd(2, B, D, h, w).to(dev()), rnd(3, B, Cv, h, w).to(dev())
        tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev()).axis_tables(Ho, Wo)
        assert ops.select_algo(q.shape, v.shape, n, K) == "union_tc"
        n0, g0 = ops.launch_count("xattn_union_tc"), ops.launch_count("xattn_generic")
        got = ops.xattn(q, k, v, n, K, rope_tables=tabs)
        got16 = ops.xattn(q, k, v, n, K, rope_tables=tabs, out_dtype=torch.bfloat16)
        assert ops.launch_count("xattn_union_tc") == n0 + 2 and ops.launch_count("xattn_generic") == g0
        want = ops.xattn(q, k, v, n, K, rope_tables=tabs, algo=_lib.ALGO_GENERIC)
        assert (got - want).abs().max().item() <= 1e-4
        assert got16.dtype == torch.bfloat16 and torch.equal(got16, got.to(torch.bfloat16))


def test_union_kernel_reads_replicated_guidance():
    """`rep` (guidance stored at the encoder resolution) through the union kernel: 2-pixel cells."""
    B, D, Cv, Hs, Ws, K, rep = 1, 256, 64, 20, 24, 5, (2, 2)
    Ho, Wo, h, w = Hs * 2, Ws * 2, 20, 24
    xs, feats = rnd(7, B, D, Hs, Ws), rnd(8, B, Cv, h, w)
    want = O.naf_forward(xs.repeat_interleave(2, 2).repeat_interleave(2, 3), feats, 4, 4, K)
    model = naf_b200.NAF(kernel_size=K).eval().to(dev())
    model.upsampler.algo = _lib.ALGO_UNION_TC
    got = model.upsample_from_guidance(xs.to(dev()), feats.to(dev()), rep=rep)
    assert (got.cpu() - want).abs().max().item() <= 5e-5
