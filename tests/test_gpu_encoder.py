"""GPU parity of the tensor-core guidance encoder (naf_conv_tc.cu) through the C ABI.

References: the same layers as torch modules on CPU in float64 (kernel-level tests) and the golden
fixtures made from the unmodified reference at the default width (tests/golden/naf256_*.npz).
Tolerances: passes=3 (split fp16, the strict-fp32 class) 2e-5 relative to the layer's max |value|;
passes=1 (operands rounded to fp16 = the TF32 class PyTorch uses by default) 4e-3 of max |value|.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as G
import naf_b200
from naf_b200 import _lib, encoder_fast, ops
from naf_b200.layers import encoder

pytestmark = pytest.mark.gpu

TOL = {3: 2e-5, 1: 4e-3}


def dev():
    return torch.device("cuda", 0)


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def tile_partials(y_nhwc):
    """Reference (B, tiles, 8, 2) partial sums of a (B,H,W,128) tensor over 16x8 tiles, float64."""
    B, H, W, C = y_nhwc.shape
    ty, tx = -(-H // 16), -(-W // 8)
    yp = torch.zeros(B, ty * 16, tx * 8, C, dtype=torch.float64)
    yp[:, :H, :W] = y_nhwc.double()
    t = yp.view(B, ty, 16, tx, 8, 8, 16)
    return torch.stack([t.sum(dim=(2, 4, 6)), (t * t).sum(dim=(2, 4, 6))], dim=-1).view(B, ty * tx, 8, 2)


SHAPES = [(1, 16, 8), (2, 37, 29), (1, 48, 40), (1, 2, 2), (3, 20, 100), (1, 70, 17), (2, 64, 64)]


@pytest.mark.parametrize("tc", [False, True], ids=["simt", "tensor-core"])
@pytest.mark.parametrize("ks", [1, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_stem_conv_and_partials(ks, shape, tc):
    """Stem conv: exact-fp32 SIMT kernel (1e-5) and the tensor-core GEMM variant (fp16-rounded operands,
    the TF32 class: 4e-3 of max |value|); both emit the GroupNorm partial sums of what they stored."""
    B, H, W = shape
    conv = torch.nn.Conv2d(3, 128, ks, padding=ks // 2, padding_mode="reflect")
    img = rnd(1, B, 3, H, W)
    want = conv.double()(img.double()).permute(0, 2, 3, 1).contiguous()
    conv = conv.float().to(dev())
    x = img.to(dev())[:, :, :, :]
    out = torch.empty(B, H, W, 128, device=dev())
    tiles = -(-H // 16) * -(-W // 8)
    part = torch.full((B, tiles, 16), float("nan"), device=dev())
    sb, sc, sy, sx = x.stride()
    fn = _lib.load().naf_enc_stem_tc_f32 if tc else _lib.load().naf_enc_stem_f32
    rc = fn(ops._ptr(x), sb, sc, sy, sx, ops._ptr(conv.weight), ops._ptr(conv.bias),
            ops._ptr(out), ops._ptr(part), B, H, W, ks, ops._stream(dev()))
    _lib.check(rc, "stem")
    tol = 4e-3 * want.abs().max().item() if tc else 1e-5
    assert (out.cpu().double() - want).abs().max().item() <= tol
    wp = tile_partials(out.cpu())                      # statistics of the values actually stored
    assert (part.cpu().double().view(B, tiles, 8, 2) - wp).abs().max().item() <= 1e-5 * max(1.0, wp.abs().max().item())


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("ks", [1, 3])
@pytest.mark.parametrize("shape", SHAPES)
def test_gn_silu_conv_kernel(ks, passes, shape):
    B, H, W = shape
    torch.manual_seed(5)
    norm = torch.nn.GroupNorm(8, 128)
    conv = torch.nn.Conv2d(128, 128, ks, padding=ks // 2, padding_mode="reflect")
    with torch.no_grad():
        norm.weight.add_(0.3 * torch.randn(128))
        norm.bias.add_(0.3 * torch.randn(128))
    y = rnd(2, B, 128, H, W) * 1.7 + 0.4
    with torch.no_grad():
        want = conv.double()(F.silu(norm.double()(y.double()))).permute(0, 2, 3, 1).contiguous()
    norm, conv = norm.float().to(dev()), conv.float().to(dev())
    y_nhwc = y.permute(0, 2, 3, 1).contiguous()
    tiles = -(-H // 16) * -(-W // 8)
    part = tile_partials(y_nhwc).float().view(B, tiles, 16).to(dev())
    yd = y_nhwc.to(dev())
    coef = torch.empty(B, 128, 2, device=dev())
    lib, st = _lib.load(), ops._stream(dev())
    _lib.check(lib.naf_enc_gn_coef_f32(ops._ptr(part), ops._ptr(norm.weight), ops._ptr(norm.bias), ops._ptr(coef),
                                       B, H, W, float(norm.eps), st), "coef")
    # coefficients against torch's GroupNorm statistics
    yg = y.double().view(B, 8, -1)
    mean, var = yg.mean(-1), yg.var(-1, unbiased=False)
    a = (1.0 / torch.sqrt(var + norm.eps)).repeat_interleave(16, 1) * norm.weight.cpu().double()[None]
    b = norm.bias.cpu().double()[None] - mean.repeat_interleave(16, 1) * a
    assert (coef.cpu().double()[..., 0] - a).abs().max().item() <= 1e-5
    assert (coef.cpu().double()[..., 1] - b).abs().max().item() <= 1e-5
    # conv into a slab of a wider tensor, with statistics of the output
    wide = torch.full((B, H, W, 256), 7.0, device=dev())
    part_out = torch.full((B, tiles, 16), float("nan"), device=dev())
    wpk = encoder_fast._packed_weight(conv)
    _lib.check(lib.naf_enc_conv_f32(ops._ptr(yd), ops._ptr(coef), ops._ptr(wpk), ops._ptr(conv.bias), ops._ptr(wide),
                                    256, 128, ops._ptr(part_out), B, H, W, ks, passes, st), "conv")
    got = wide.cpu().double()
    assert (got[..., :128] == 7.0).all()                       # the other slab is untouched
    scale = want.abs().max().item()
    err = (got[..., 128:] - want).abs().max().item()
    assert err <= TOL[passes] * scale, (ks, passes, shape, err, scale)
    # partial sums: kernels tile differently (16x8 or 32x8), what is pinned is the per-(image, group) total
    wp = tile_partials(got[..., 128:].float()).sum(dim=1)
    gp = part_out.cpu().double().view(B, tiles, 8, 2).sum(dim=1)
    assert (gp - wp).abs().max().item() <= 1e-5 * max(1.0, wp.abs().max().item())


@pytest.mark.parametrize("passes", [3, 1])
@pytest.mark.parametrize("ks", [1, 3])
def test_whole_branch_matches_torch_modules(ks, passes):
    torch.manual_seed(3)
    seq = encoder(3, 128, kernel_size=ks, ks_res=ks, num_layers=2).eval()
    img = rnd(9, 2, 3, 40, 52)
    with torch.no_grad():
        want = seq.double()(img.double()).permute(0, 2, 3, 1)
    seq = seq.float().to(dev())
    assert encoder_fast.tc_supported(seq)
    got = encoder_fast.forward_tc(seq, img.to(dev()), passes=passes)
    scale = want.abs().max().item()
    err = (got.cpu().double() - want).abs().max().item()
    assert err <= 4 * TOL[passes] * scale, (ks, passes, err, scale)


@pytest.mark.parametrize("name", G.names("naf256_"))
def test_default_width_module_matches_reference_golden(name):
    """Whole NAF.forward at the default width: tensor-core encoder (strict class, allow_tf32=False
    -> 3 passes) + RoPE/key-pool + attention kernels vs the unmodified reference."""
    c = G.default_case(name)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = naf_b200.NAF(kernel_size=7).eval()
        m.load_state_dict(G.default_state(), strict=True)
        m = m.to(dev())
        n0 = ops.launch_count("enc_conv")
        out = m(c["image"].to(dev()), c["features"].to(dev()), c["output_size"])
        assert ops.launch_count("enc_conv") > n0   # the fused tensor-core encoder must have run
        assert out.shape == c["out"].shape
        assert (out.cpu() - c["out"]).abs().max().item() <= 1e-4, name
        q = m.image_encoder(c["image"].to(dev()), c["output_size"])
        assert (q.cpu()[:, ::c["qstep"]] - c["queries_sub"]).abs().max().item() <= 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_default_width_module_tf32_class_error_is_bounded():
    """With PyTorch's default flags (allow_tf32=True) the encoder runs the 1-pass kernels; the whole
    forward must still meet BASELINE.json's 1e-3 bar against the fp32 reference golden."""
    c = G.default_case("naf256_same_res")
    assert torch.backends.cudnn.allow_tf32
    m = naf_b200.NAF(kernel_size=7).eval()
    m.load_state_dict(G.default_state(), strict=True)
    m = m.to(dev())
    out = m(c["image"].to(dev()), c["features"].to(dev()), c["output_size"])
    err = (out.cpu() - c["out"]).abs().max().item()
    print("tf32-class whole-module max|err| vs reference golden =", err)
    assert err <= 1e-3


def test_bench_path_tf32_class_meets_the_bar_on_a_c2_shaped_case():
    """The path bench.py times (default flags: 1-pass fp16-operand encoder) against the strict class
    (3-pass, held to 1e-4 of the reference goldens above) on ONE C2-shaped image: guidance 448 -> target
    896, C=768 features 32x32, K=7, random-init weights.  Bar: BASELINE.json's 1e-3 max-abs minus the
    strict class's own 1e-4 allowance."""
    torch.manual_seed(7)
    m = naf_b200.NAF(kernel_size=7).eval().to(dev())
    g = torch.Generator(device="cpu").manual_seed(3)
    img = torch.randn(1, 3, 448, 448, generator=g).to(dev())
    ft = torch.randn(1, 768, 32, 32, generator=g).to(dev())
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        fast = m(img, ft, (896, 896))
        torch.backends.cudnn.allow_tf32 = False
        strict = m(img, ft, (896, 896))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    err = (fast - strict).abs().max().item()
    print("C2-shaped: tf32-class vs strict-class whole-forward max|diff| =", err)
    assert err <= 9e-4


# ------------------------------------------------------------------ fp16 activations between the layers (TF32 class)
@pytest.mark.parametrize("io", [("f16", "f16"), ("f16", "f32"), ("f32", "f16")])
@pytest.mark.parametrize("ks", [1, 3])
@pytest.mark.parametrize("shape", [(1, 16, 8), (2, 40, 52), (1, 37, 29)])
def test_gn_silu_conv_kernel_fp16_activations(ks, io, shape):
    """naf_enc_conv_ex with fp16 input and / or output activations (1-pass class): same bar as the fp32-I/O 1-pass
    kernel (4e-3 of the layer's max |value|); the partial sums are those of the fp32 accumulators."""
    B, H, W = shape
    torch.manual_seed(6)
    norm = torch.nn.GroupNorm(8, 128)
    conv = torch.nn.Conv2d(128, 128, ks, padding=ks // 2, padding_mode="reflect")
    with torch.no_grad():
        norm.weight.add_(0.3 * torch.randn(128))
        norm.bias.add_(0.3 * torch.randn(128))
    y = rnd(4, B, 128, H, W) * 1.7 + 0.4
    in16, out16 = io[0] == "f16", io[1] == "f16"
    y_nhwc = y.permute(0, 2, 3, 1).contiguous()
    y_in = y_nhwc.half() if in16 else y_nhwc
    y_seen = y_in.float().permute(0, 3, 1, 2)          # what the kernel reads, as NCHW
    with torch.no_grad():
        want = conv.double()(F.silu(norm.double()(y_seen.double()))).permute(0, 2, 3, 1).contiguous()
    norm, conv = norm.float().to(dev()), conv.float().to(dev())
    tiles = -(-H // 16) * -(-W // 8)
    part = tile_partials(y_in.float()).float().view(B, tiles, 16).to(dev())
    coef = torch.empty(B, 128, 2, device=dev())
    lib, st = _lib.load(), ops._stream(dev())
    _lib.check(lib.naf_enc_gn_coef_f32(ops._ptr(part), ops._ptr(norm.weight), ops._ptr(norm.bias), ops._ptr(coef),
                                       B, H, W, float(norm.eps), st), "coef")
    yd = y_in.to(dev())
    out = torch.full((B, H, W, 128), 7.0, device=dev(), dtype=torch.float16 if out16 else torch.float32)
    part_out = torch.full((B, tiles, 16), float("nan"), device=dev())
    wpk = encoder_fast._packed_weight(conv)
    code = {"f16": _lib.DTYPE_F16, "f32": _lib.DTYPE_F32}
    _lib.check(lib.naf_enc_conv_ex(ops._ptr(yd), ops._ptr(coef), ops._ptr(wpk), ops._ptr(conv.bias), ops._ptr(out),
                                   128, 0, ops._ptr(part_out), B, H, W, ks, 1, code[io[0]], code[io[1]], st), "conv")
    got = out.cpu().double()
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= 4e-3 * scale, (ks, io, shape, err, scale)
    wp = tile_partials(want.float()).sum(dim=1)          # statistics of the (un-rounded) result
    gp = part_out.cpu().double().view(B, tiles, 8, 2).sum(dim=1)
    assert (gp - wp).abs().max().item() <= 4e-3 * max(1.0, wp.abs().max().item())
    with pytest.raises(NotImplementedError):           # fp16 activations are a TF32-class feature
        _lib.check(lib.naf_enc_conv_ex(ops._ptr(yd), ops._ptr(coef), ops._ptr(wpk), ops._ptr(conv.bias), ops._ptr(out),
                                       128, 0, ops._ptr(part_out), B, H, W, ks, 3, _lib.DTYPE_F16, _lib.DTYPE_F16, st), "conv")


@pytest.mark.parametrize("ks,tc", [(1, 0), (3, 0), (3, 1), (1, 1)])
def test_stem_fp16_output(ks, tc):
    B, H, W = 2, 40, 52
    torch.manual_seed(2)
    conv = torch.nn.Conv2d(3, 128, ks, padding=ks // 2, padding_mode="reflect")
    x = rnd(3, B, 3, H, W)
    with torch.no_grad():
        want = conv.double()(x.double()).permute(0, 2, 3, 1).contiguous()
    conv = conv.float().to(dev())
    xd = x.to(dev())
    tiles = -(-H // 16) * -(-W // 8)
    out = torch.empty((B, H, W, 128), device=dev(), dtype=torch.float16)
    part = torch.empty((B, tiles, 16), device=dev())
    sb, sc, sy, sx = xd.stride()
    rc = _lib.load().naf_enc_stem_ex(ops._ptr(xd), sb, sc, sy, sx, ops._ptr(conv.weight), ops._ptr(conv.bias),
                                     ops._ptr(out), ops._ptr(part), B, H, W, ks, tc, _lib.DTYPE_F16, ops._stream(dev()))
    _lib.check(rc, "stem")
    scale = want.abs().max().item()
    assert (out.cpu().double() - want).abs().max().item() <= (4e-3 if tc else 1e-3) * scale
    wp = tile_partials(want.float())
    assert (part.cpu().double().view(B, tiles, 8, 2) - wp).abs().max().item() <= 4e-3 * max(1.0, wp.abs().max().item())


@pytest.mark.parametrize("ks", [1, 3])
def test_whole_branch_fp32_activations_still_available(ks):
    """The TF32-class branch with fp32 activations between the layers (round 1's data flow) stays selectable and
    agrees with the default fp16-activation flow to the class tolerance."""
    torch.manual_seed(3)
    seq = encoder(3, 128, kernel_size=ks, ks_res=ks, num_layers=2).eval().to(dev())
    img = rnd(9, 2, 3, 40, 52).to(dev())
    a = encoder_fast.forward_tc(seq, img, passes=1, act_dtype=torch.float32)
    b = encoder_fast.forward_tc(seq, img, passes=1)
    scale = a.abs().max().item()
    assert (a - b).abs().max().item() <= 8e-3 * scale
    with pytest.raises(ValueError):
        encoder_fast.forward_tc(seq, img, passes=3, act_dtype=torch.float16)
