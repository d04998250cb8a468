"""Forward of the guidance conv encoder on our kernels.

The reference's `encoder()` (src/layers/convolutions.py:70-95) is a stem conv followed by EncBlocks
(GroupNorm -> SiLU -> conv -> GroupNorm -> SiLU -> conv, reflect padding, no residual).  Two paths:

* `forward_tc` (the one `NAF.forward` takes for the reference-default 128-channel branches): the whole
  stack on tcgen05 kernels of libnaf_b200.so -- stem conv, per EncBlock half one GroupNorm-coefficient
  kernel and ONE fused GroupNorm+SiLU+reflect-pad+conv+bias kernel (naf_conv_tc.cu).  No library
  convolution is called.
* `forward` (other channel widths): cuDNN convolutions (channels_last, no bias) with everything
  between them replaced by two of our kernels per GroupNorm (`naf_gn_stats_f32`,
  `naf_gn_silu_apply_f32`).

Precision classes follow `torch.backends.cudnn.allow_tf32` (see `conv_passes`): operands rounded to
fp16 (10-bit mantissa like TF32, but fp16's 5-bit exponent: |x| <= 65504, and values below 6e-5 lose
bits -- the GroupNorm-normalised activations, the weights and an ImageNet-normalised image (the
reference's calling convention, evaluation/eval_seg_probing.py:98) sit far inside that range; raw
un-normalised 16-bit imagery does not and must use the strict class), or split-fp16 operands with
fp32-class accuracy and the exact fp32 SIMT stem (`allow_tf32 = False`, no range restriction beyond
fp32's for the image).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops
from .layers.convolutions import EncBlock


def _conv_ok(conv) -> bool:
    k = conv.kernel_size
    return (isinstance(conv, nn.Conv2d) and k[0] == k[1] and k[0] in (1, 3) and conv.stride == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1
            and (k[0] == 1 or conv.padding_mode == "reflect") and conv.padding == (k[0] // 2, k[0] // 2))


def supported(seq) -> bool:
    """True if `seq` has exactly the structure the fast path implements."""
    if not isinstance(seq, nn.Sequential) or len(seq) < 1 or not _conv_ok(seq[0]):
        return False
    for blk in list(seq)[1:]:
        if not isinstance(blk, EncBlock) or blk.residual or blk.use_conv_shortcut:
            return False
        if not isinstance(blk.activation_fn, nn.SiLU):
            return False
        for norm, conv in ((blk.norm1, blk.conv1), (blk.norm2, blk.conv2)):
            if not isinstance(norm, nn.GroupNorm) or not norm.affine or not _conv_ok(conv):
                return False
            Cn, G = norm.num_channels, norm.num_groups
            if Cn % 4 or (Cn // 4) % G or 256 % (Cn // 4) or (Cn // G) % 4 or G > 64 or conv.in_channels != Cn:
                return False
    return True


def _weight_cl(conv):
    """channels_last copy of the conv weight, cached on the module."""
    w = conv.weight
    key = (w.data_ptr(), w._version, w.device)
    cache = getattr(conv, "_naf_wcl", None)
    if cache is None or cache[0] != key:
        cache = (key, w.detach().contiguous(memory_format=torch.channels_last))
        conv._naf_wcl = cache
    return cache[1]


def _conv_nhwc(x_nhwc, conv):
    """cuDNN convolution on an ALREADY padded pixel-major tensor, no bias; returns pixel-major."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), _weight_cl(conv), None)
    y = y.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


def groupnorm_silu(y, prev_bias, norm, pad: int):
    """silu(GroupNorm(y + prev_bias)) written with a reflect border of `pad`; y is (B,H,W,C)."""
    B, H, W, Cn = y.shape
    dev = y.device
    G = norm.num_groups
    sums = torch.empty((B, G, 2), device=dev, dtype=torch.float64)
    out = torch.empty((B, H + 2 * pad, W + 2 * pad, Cn), device=dev, dtype=torch.float32)
    lib = _lib.load()
    st = ops._stream(dev)
    with torch.cuda.device(dev):
        rc = lib.naf_gn_stats_f32(ops._ptr(y), ops._ptr(prev_bias), ops._ptr(sums), B, H * W, Cn, G, st)
        _lib.check(rc, "naf_gn_stats_f32")
        rc = lib.naf_gn_silu_apply_f32(ops._ptr(y), ops._ptr(prev_bias), ops._ptr(norm.weight), ops._ptr(norm.bias),
                                       ops._ptr(sums), ops._ptr(out), B, H, W, Cn, G, float(norm.eps), pad, st)
        _lib.check(rc, "naf_gn_silu_apply_f32")
    return out


# ------------------------------------------------------------------------ tensor-core encoder
TILE_H, TILE_W = 16, 8   # pixel tile of the conv kernels (naf_conv_tc.cu)


def tc_supported(seq) -> bool:
    """The fused tensor-core stack handles exactly the NAF default branch: 3 -> 128 stem, EncBlocks
    of 128 channels with GroupNorm(8), SiLU, bias, reflect padding, one kernel size (1 or 3)."""
    if not supported(seq) or not _lib.load().naf_has_tensor_path():
        return False
    stem = seq[0]
    k = stem.kernel_size[0]
    if stem.in_channels != 3 or stem.out_channels != 128 or stem.bias is None:
        return False
    for blk in list(seq)[1:]:
        for norm, conv in ((blk.norm1, blk.conv1), (blk.norm2, blk.conv2)):
            if (norm.num_channels != 128 or norm.num_groups != 8 or conv.out_channels != 128
                    or conv.kernel_size[0] != k or conv.bias is None):
                return False
    return True


def conv_passes() -> int:
    """1 = fp16-rounded operands (the precision class of PyTorch's default TF32 convolutions),
    3 = split-fp16 with fp32-class accuracy; follows `torch.backends.cudnn.allow_tf32` like the
    reference's own convolutions do."""
    return 1 if torch.backends.cudnn.allow_tf32 else 3


def _packed_weight(conv):
    """fp16 hi/lo operand image of a 128x128xKxK conv weight (naf_enc_conv_pack_f32), cached."""
    w = conv.weight
    key = (w.data_ptr(), w._version, w.device)
    cache = getattr(conv, "_naf_wtc", None)
    if cache is None or cache[0] != key:
        k = conv.kernel_size[0]
        packed = torch.empty(k * k * 65536, device=w.device, dtype=torch.uint8)
        wc = w.detach().float().contiguous()
        with torch.cuda.device(w.device):
            rc = _lib.load().naf_enc_conv_pack_f32(ops._ptr(wc), ops._ptr(packed), k, ops._stream(w.device))
        _lib.check(rc, "naf_enc_conv_pack_f32")
        cache = (key, packed)
        conv._naf_wtc = cache
    return cache[1]


def forward_tc(seq, image, out=None, ch_off=0, passes=None, act_dtype=None):
    """`seq(image)` on our kernels only: stem conv, then per EncBlock half one tiny coefficient
    kernel and ONE fused GroupNorm+SiLU+conv kernel.  Returns the pixel-major (B,H,W,C_total) fp32
    tensor `out` whose channels [ch_off, ch_off+128) hold the result (bias included); `out` is
    allocated (128 channels) when not given.

    `act_dtype`: element type of the activations BETWEEN the layers.  Default: fp16 in the TF32 class
    (passes = 1: the next layer rounds its input operand to fp16 anyway, so every layer moves half the
    HBM bytes; GroupNorm statistics still come from the fp32 accumulators), fp32 in the strict class."""
    lib = _lib.load()
    dev = image.device
    x = image if image.dtype == torch.float32 else image.float()
    B, _, H, W = x.shape
    passes = conv_passes() if passes is None else int(passes)
    if act_dtype is None:
        act_dtype = torch.float16 if passes == 1 else torch.float32
    if act_dtype not in (torch.float16, torch.float32) or (act_dtype == torch.float16 and passes != 1):
        raise ValueError("forward_tc: fp16 activations belong to the 1-pass (TF32) class only")
    act_code = _lib.DTYPE_F16 if act_dtype == torch.float16 else _lib.DTYPE_F32
    stem = seq[0]
    k = stem.kernel_size[0]
    tiles = -(-H // TILE_H) * -(-W // TILE_W)
    st = ops._stream(dev)
    layers = [(n, c) for blk in list(seq)[1:] for n, c in ((blk.norm1, blk.conv1), (blk.norm2, blk.conv2))]
    if out is None:
        out = torch.empty((B, H, W, 128), device=dev, dtype=torch.float32)
        ch_off = 0
    assert out.is_contiguous() and tuple(out.shape[:3]) == (B, H, W) and out.dtype == torch.float32
    pix_stride = out.shape[-1]
    part = torch.empty((B, tiles, 16), device=dev, dtype=torch.float32)
    coef = torch.empty((B, 128, 2), device=dev, dtype=torch.float32)
    bufs = [torch.empty((B, H, W, 128), device=dev, dtype=act_dtype) for _ in range(min(2, len(layers)))]
    sb, sc, sy, sx = x.stride()
    with torch.cuda.device(dev):
        last = len(layers) == 0
        y = out if last else bufs[0]
        # TF32 class, 3x3: the stem as a tensor-core GEMM (0.24 vs 0.40 ms at C2); the 1x1 stem (3 MACs per
        # output) and the strict class keep the exact-fp32 SIMT kernel (0.15 ms vs 0.21 ms on the GEMM path)
        rc = lib.naf_enc_stem_ex(ops._ptr(x), sb, sc, sy, sx, ops._ptr(stem.weight), ops._ptr(stem.bias),
                                 ops._ptr(y), ops._ptr(None if last else part), B, H, W, k,
                                 1 if (passes == 1 and k == 3) else 0, _lib.DTYPE_F32 if last else act_code, st)
        _lib.check(rc, "naf_enc_stem_ex")
        if last and (pix_stride != 128 or ch_off):
            raise NotImplementedError("stem-only encoder into a slab")
        for i, (norm, conv) in enumerate(layers):
            last = i == len(layers) - 1
            rc = lib.naf_enc_gn_coef_f32(ops._ptr(part), ops._ptr(norm.weight), ops._ptr(norm.bias),
                                         ops._ptr(coef), B, H, W, float(norm.eps), st)
            _lib.check(rc, "naf_enc_gn_coef_f32")
            dst = out if last else bufs[(i + 1) & 1]
            rc = lib.naf_enc_conv_ex(ops._ptr(y), ops._ptr(coef), ops._ptr(_packed_weight(conv)),
                                     ops._ptr(conv.bias), ops._ptr(dst), pix_stride if last else 128,
                                     ch_off if last else 0, ops._ptr(None if last else part), B, H, W, k,
                                     passes, act_code, _lib.DTYPE_F32 if last else act_code, st)
            _lib.check(rc, "naf_enc_conv_ex")
            y = dst
    return out


def forward(seq, image):
    """Pixel-major (B,H,W,C) output of `seq(image)` WITHOUT the bias of the last convolution, and
    that bias (or None): the caller folds it into the concat pass."""
    stem = seq[0]
    k = stem.kernel_size[0]
    x = image.float()
    if k == 3:
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    x = x.contiguous(memory_format=torch.channels_last)
    y = F.conv2d(x, _weight_cl(stem), None).permute(0, 2, 3, 1)
    y = y if y.is_contiguous() else y.contiguous()
    bias = stem.bias
    for blk in list(seq)[1:]:
        for norm, conv in ((blk.norm1, blk.conv1), (blk.norm2, blk.conv2)):
            a = groupnorm_silu(y, bias, norm, conv.kernel_size[0] // 2)
            y = _conv_nhwc(a, conv)
            bias = conv.bias
    return y, bias
