"""Functional layer over the C ABI: torch tensors in, torch tensors out, kernels from
libnaf_b200.so in between.  PyTorch is used for device memory, streams and (tiny) table maths
only -- every pass over the big tensors is one of our CUDA kernels.

Layout convention: public tensors are NCHW-*shaped* like the reference's; internally everything
is "pixel-major" (channels innermost).  A channels_last NCHW tensor already IS pixel-major, so
`as_pixel_major` is free for it; anything else goes through the packing kernel once.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib, taps


def launch_count(prefix: str = "") -> int:
    """Kernels launched by libnaf_b200.so since load whose family starts with `prefix` (native
    counter, naf_launch_count); "" = all."""
    return int(_lib.load().naf_launch_count(prefix.encode()))


# ------------------------------------------------------------------------------------ helpers
def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "naf_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback "
                f"(got a tensor on {t.device})")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def _no_grad_guard(*tensors: torch.Tensor) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(
            "naf_b200.ops functions are the forward pass only: call them under torch.no_grad() / "
            "torch.inference_mode(), detach the inputs, or go through the differentiable modules "
            "(naf_b200.NAF / CrossAttention / RoPE, naf_b200.autograd)")


def _stream(dev: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def is_pixel_major(t: torch.Tensor) -> bool:
    """NCHW-shaped tensor whose channel stride is 1 and whose pixel strides are 16 B friendly."""
    if t.dim() != 4:
        return False
    sb, sc, sy, sx = t.stride()
    Cn = t.shape[1]
    if Cn > 1 and sc != 1:
        return False
    return sx >= Cn and sx % 4 == 0 and sy % 4 == 0 and sb % 4 == 0 and t.data_ptr() % 16 == 0


def pack_nhwc(t: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) fp32, any strides -> contiguous (B,H,W,C) via naf_pack_nhwc_f32."""
    dev = _require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    B, Cn, H, W = t.shape
    out = torch.empty((B, H, W, Cn), device=dev, dtype=torch.float32)
    sb, sc, sy, sx = t.stride()
    with torch.cuda.device(dev):
        rc = _lib.load().naf_pack_nhwc_f32(_ptr(t), _ptr(out), B, Cn, H, W, sb, sc, sy, sx, _stream(dev))
    _lib.check(rc, "naf_pack_nhwc_f32")
    return out


def pack_concat_nhwc(parts) -> torch.Tensor:
    """torch.cat(parts, dim=1) + pixel-major packing in one pass per part (naf_pack_nhwc_slab_f32):
    parts are (B,Ci,H,W) fp32 tensors of any strides; returns the NCHW-shaped pixel-major view of
    a contiguous (B,H,W,sum Ci) buffer."""
    dev = _require_cuda(*parts)
    B, _, H, W = parts[0].shape
    total = sum(int(t.shape[1]) for t in parts)
    out = torch.empty((B, H, W, total), device=dev, dtype=torch.float32)
    off = 0
    with torch.cuda.device(dev):
        for t in parts:
            if t.dtype != torch.float32:
                t = t.float()
            if tuple(t.shape[0:1] + t.shape[2:]) != (B, H, W):
                raise ValueError("pack_concat_nhwc: parts disagree on (B, H, W)")
            sb, sc, sy, sx = t.stride()
            rc = _lib.load().naf_pack_nhwc_slab_f32(_ptr(t), _ptr(out), B, t.shape[1], H, W, sb, sc, sy, sx,
                                                    total, off, _stream(dev))
            _lib.check(rc, "naf_pack_nhwc_slab_f32")
            off += int(t.shape[1])
    return out.permute(0, 3, 1, 2)


def concat_bias_nhwc(a, bias_a, b, bias_b) -> torch.Tensor:
    """cat([a + bias_a, b + bias_b], channels) of two pixel-major (B,H,W,C*) tensors in one pass;
    returns the NCHW-shaped pixel-major view."""
    dev = _require_cuda(a, b)
    B, H, W, Ca = a.shape
    Cb = b.shape[-1]
    assert a.is_contiguous() and b.is_contiguous() and tuple(b.shape[:3]) == (B, H, W)
    out = torch.empty((B, H, W, Ca + Cb), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        rc = _lib.load().naf_concat_bias_nhwc_f32(_ptr(a), _ptr(bias_a), Ca, _ptr(b), _ptr(bias_b), Cb,
                                                  _ptr(out), B * H * W, _stream(dev))
    _lib.check(rc, "naf_concat_bias_nhwc_f32")
    return out.permute(0, 3, 1, 2)


def as_pixel_major(t: torch.Tensor) -> torch.Tensor:
    """NCHW-shaped view whose storage is pixel-major (no copy when it already is)."""
    if t.dtype != torch.float32:
        t = t.float()
    if is_pixel_major(t):
        return t
    return pack_nhwc(t).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------- rope tables
def rope_tables_from_coords(cy: torch.Tensor, cx: torch.Tensor, periods: torch.Tensor):
    """Per-axis cos/sin tables (cos_y, sin_y: (H,P); cos_x, sin_x: (W,P)), P = len(periods), from the
    per-axis coordinates.  Same torch ops as the reference's `RoPE.rotate` (src/layers/rope.py:139-146)
    so the entries are bit-identical to those of its (H*W, D_head) cos/sin tensors on the same device."""
    ang_y = 2 * math.pi * cy[:, None] / periods[None, :]
    ang_x = 2 * math.pi * cx[:, None] / periods[None, :]
    return (torch.cos(ang_y).contiguous(), torch.sin(ang_y).contiguous(),
            torch.cos(ang_x).contiguous(), torch.sin(ang_x).contiguous())


def rope_axis_tables(H: int, W: int, periods: torch.Tensor):
    """Tables for the default ("separate", eval) coordinates (src/layers/rope.py:99-105)."""
    dev, dt = periods.device, periods.dtype
    cy = 2.0 * (torch.arange(0.5, H, device=dev, dtype=dt) / H) - 1.0
    cx = 2.0 * (torch.arange(0.5, W, device=dev, dtype=dt) / W) - 1.0
    return rope_tables_from_coords(cy, cx, periods)


# -------------------------------------------------------------------------- rope + key pooling
def rope_kpool(x: torch.Tensor, tables, rope_heads: int, pooled_hw=None, want_q: bool = False,
               rep=(1, 1)):
    """x (B,D,Ho,Wo) -> (k (B,D,h,w) or None, q (B,D,Ho,Wo) or None), both pixel-major views.

    tables = (cos_y, sin_y, cos_x, sin_x) or None (x already rotated).
    rep = (ry, rx): x is a SOURCE map (B,D,Ho/ry,Wo/rx) each pixel of which stands for an
    ry x rx block of target pixels (the reference's adaptive_avg_pool2d up-replication)."""
    dev = _require_cuda(x)
    x = as_pixel_major(x)
    B, D, Ho, Wo = x.shape
    Ho, Wo = Ho * int(rep[0]), Wo * int(rep[1])
    k = q = None
    h = w = 0
    if pooled_hw is not None:
        h, w = int(pooled_hw[0]), int(pooled_hw[1])
        k = torch.empty((B, h, w, D), device=dev, dtype=torch.float32)
    if want_q:
        q = torch.empty((B, Ho, Wo, D), device=dev, dtype=torch.float32)
    if k is None and q is None:
        raise ValueError("rope_kpool: nothing to compute")
    p = _lib.KPoolParams()
    p.x, p.k_out, p.q_out = _ptr(x), _ptr(k), _ptr(q)
    if tables is not None:
        for name, t in zip(("cos_y", "sin_y", "cos_x", "sin_x"), tables):
            assert t.is_contiguous() and t.dtype == torch.float32 and t.device == dev
            setattr(p, name, _ptr(t))
        if tables[0].shape[0] != Ho or tables[2].shape[0] != Wo:
            raise ValueError("rope tables do not match the map size")
    p.B, p.D, p.Ho, p.Wo, p.h, p.w = B, D, Ho, Wo, h, w
    p.rope_heads = int(rope_heads)
    p.x_stride_b, _, p.x_stride_y, p.x_stride_x = x.stride()
    p.rep_y, p.rep_x = int(rep[0]), int(rep[1])
    with torch.cuda.device(dev):
        rc = _lib.load().naf_rope_kpool_f32(C.byref(p), _stream(dev))
    _lib.check(rc, "naf_rope_kpool_f32")
    return (None if k is None else k.permute(0, 3, 1, 2),
            None if q is None else q.permute(0, 3, 1, 2))


# ---------------------------------------------------------------------------------- attention
_tap_cache: dict = {}


def device_tap_tables(Ho, Wo, h, w, K, dev):
    """int32 device tables (row_tap, col_tap) or (None, None) for integer ratios; cached."""
    K = taps._pair(K)
    key = (Ho, Wo, h, w, K, str(dev))
    hit = _tap_cache.get(key)
    if hit is None:
        rt, ct = taps.tap_tables(Ho, Wo, h, w, K)
        if rt is None:
            hit = (None, None)
        else:
            hit = (torch.from_numpy(rt.copy()).to(dev), torch.from_numpy(ct.copy()).to(dev))
        if len(_tap_cache) > 64:
            _tap_cache.clear()
        _tap_cache[key] = hit
    return hit


def _fill_xattn(q, k, v, out, scores, heads, K, scale, tap_tabs, rope_tabs, algo, rep=(1, 1)):
    """naf_xattn_params for these tensors (out may be fp32 or bf16)."""
    B, D, Ho, Wo = q.shape
    Ho, Wo = Ho * int(rep[0]), Wo * int(rep[1])
    _, Cn, h, w = v.shape
    p = _lib.XAttnParams()
    p.q, p.k, p.v, p.out, p.scores = _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(scores)
    p.row_tap, p.col_tap = _ptr(tap_tabs[0]), _ptr(tap_tabs[1])
    if rope_tabs is not None:
        p.cos_y, p.sin_y, p.cos_x, p.sin_x = (_ptr(t) for t in rope_tabs)
    kh, kw = taps._pair(K)
    p.B, p.D, p.C, p.heads, p.Ho, p.Wo, p.h, p.w, p.K = B, D, Cn, heads, Ho, Wo, h, w, kh
    p.Kw = 0 if kw == kh else kw      # rectangular window (NATTEN kernel_size=(kh, kw)); 0 = square
    p.scale = float(scale)
    p.q_stride_b, _, p.q_stride_y, p.q_stride_x = q.stride()
    p.algo = int(algo)
    p.rep_y, p.rep_x = int(rep[0]), int(rep[1])
    p.out_dtype = _lib.DTYPE_BF16 if out.dtype == torch.bfloat16 else _lib.DTYPE_F32
    p.q_dtype, p.k_dtype, p.v_dtype = (_lib.DTYPE_BF16 if t.dtype == torch.bfloat16 else _lib.DTYPE_F32 for t in (q, k, v))
    return p


def _native_pixel_major(t: torch.Tensor, contiguous: bool) -> torch.Tensor:
    """bf16 tensor as an NCHW-shaped view of pixel-major storage WITHOUT widening it (layout changes, when
    needed at all, are torch's channels_last copy of the bf16 data: plumbing on the small maps)."""
    sb, sc, sy, sx = t.stride()
    Cn = t.shape[1]
    ok = (Cn == 1 or sc == 1) and sx >= Cn and sx % 16 == 0 and sy % 16 == 0 and sb % 16 == 0 and t.data_ptr() % 32 == 0
    if ok and (not contiguous or t.permute(0, 2, 3, 1).is_contiguous()):
        return t
    return t.contiguous(memory_format=torch.channels_last)


def _launch_xattn(p, dev) -> None:
    """The one place the attention kernel is enqueued (bench.py wraps this function with CUDA
    events for its live roofline measurement; the product itself carries no instrumentation)."""
    with torch.cuda.device(dev):
        rc = _lib.load().naf_xattn_fwd_f32(C.byref(p), _stream(dev))
    _lib.check(rc, "naf_xattn_fwd_f32")


def xattn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, kernel_size,
          scale: Optional[float] = None, rope_tables=None, return_scores: bool = False,
          algo: int = _lib.ALGO_AUTO, rep=(1, 1), out_dtype: torch.dtype = torch.float32):
    """Cross-scale neighbourhood attention.  q (B,D,Ho,Wo), k (B,D,h,w), v (B,C,h,w), all
    NCHW-shaped; returns out (B,C,Ho,Wo) as a permuted view of pixel-major storage (exactly what
    the reference returns, src/layers/attentions.py:75) and optionally the scaled pre-softmax
    scores (B,heads,Ho,Wo,K*K).  kernel_size: an odd int, or NATTEN's pair (kh, kw) -- rectangular windows
    run on the generic kernel.

    out_dtype: torch.float32, or torch.bfloat16 (what the reference returns under bf16 autocast;
    the arithmetic stays fp32, only the final store is rounded).
    rope_tables: if given, q is the UN-rotated map and RoPE is applied inside the kernel.
    rep: q is a replicated source map (see rope_kpool); the target size is q's size times rep."""
    dev = _require_cuda(q, k, v)
    _no_grad_guard(q, k, v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be 4-D (B, C, H, W)")
    B, D, Ho, Wo = q.shape
    Ho, Wo = Ho * int(rep[0]), Wo * int(rep[1])
    if k.shape[0] != B or v.shape[0] != B or k.shape[1] != D or k.shape[-2:] != v.shape[-2:]:
        raise ValueError(f"inconsistent shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)}")
    assert D % heads == 0, "dim must be divisible by num_heads"
    Cn, h, w = v.shape[1:]
    if Cn % heads != 0:
        raise ValueError(f"value channels ({Cn}) must be divisible by num_heads ({heads})")
    K = taps._pair(kernel_size)
    if scale is None:
        scale = (D // heads) ** -0.5
    tap_tabs = device_tap_tables(Ho, Wo, h, w, K, dev)  # validates the window too
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise NotImplementedError(f"xattn: out_dtype {out_dtype} (float32 or bfloat16)")
    out = torch.empty((B, Ho, Wo, Cn), device=dev, dtype=out_dtype)
    scores = (torch.empty((B, heads, Ho, Wo, K[0] * K[1]), device=dev, dtype=torch.float32)
              if return_scores else None)
    p = None
    if any(t.dtype == torch.bfloat16 for t in (q, k, v)) and algo in (_lib.ALGO_AUTO, _lib.ALGO_CELL_TMA):
        # bf16 inputs (the reference under autocast, train.py:120): the TMA kernel reads them as they are
        qn, kn, vn = (_native_pixel_major(t, c) if t.dtype == torch.bfloat16 else None
                      for t, c in ((q, False), (k, True), (v, True)))
        qn = qn if qn is not None else as_pixel_major(q)
        kn = kn if kn is not None else _contig_pixel_major(k)
        vn = vn if vn is not None else _contig_pixel_major(v)
        pn = _fill_xattn(qn, kn, vn, out, scores, heads, K, scale, tap_tabs, rope_tables, algo, rep)
        pn.workspace, pn.workspace_bytes = 256, 1 << 62      # "if it had its workspace"
        if _lib.load().naf_xattn_select_algo(C.byref(pn)) == _lib.ALGO_CELL_TMA:
            pn.workspace, pn.workspace_bytes = 0, 0
            p, q, k, v = pn, qn, kn, vn
    if p is None:
        q = as_pixel_major(q)           # (widens anything that is not fp32)
        k = _contig_pixel_major(k)      # k and v must be fully contiguous pixel-major
        v = _contig_pixel_major(v)
        p = _fill_xattn(q, k, v, out, scores, heads, K, scale, tap_tabs, rope_tables, algo, rep)
    # scratch for the TMA kernel's fp16 K / V planes (the C side never allocates); freed to the caching
    # allocator when this function returns -- stream-ordered, so the enqueued kernels still own it
    ws_bytes = int(_lib.load().naf_xattn_workspace_bytes(C.byref(p)))
    if ws_bytes:
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        p.workspace, p.workspace_bytes = _ptr(ws), ws_bytes
    _launch_xattn(p, dev)
    res = out.permute(0, 3, 1, 2)
    return (res, scores) if return_scores else res


# ----------------------------------------------------------------------------------- backward
def _contig_pixel_major(t: torch.Tensor) -> torch.Tensor:
    """NCHW-shaped fp32 view whose (B,H,W,C) storage is fully contiguous."""
    t = as_pixel_major(t)
    if not t.permute(0, 2, 3, 1).is_contiguous():
        t = pack_nhwc(t).permute(0, 3, 1, 2)
    return t


def xattn_bwd(q, k, v, dout, heads: int, kernel_size, scale: Optional[float] = None, rope_tables=None,
              algo: int = _lib.ALGO_AUTO, rep=(1, 1)):
    """Gradients of `xattn` (naf_xattn_bwd_f32).  q, k, v, rope_tables, rep, scale as given to the
    forward; dout (B,C,Ho,Wo) is dL/dout.  Returns (dq, dk, dv) as NCHW-shaped pixel-major views:
    dq (B,D,Ho,Wo) is the gradient w.r.t. the ROTATED queries (w.r.t. q itself without rope tables),
    dk (B,D,h,w), dv (B,C,h,w)."""
    dev = _require_cuda(q, k, v, dout)
    B, D, Ho, Wo = q.shape
    Ho, Wo = Ho * int(rep[0]), Wo * int(rep[1])
    _, Cn, h, w = v.shape
    K = taps._pair(kernel_size)
    if tuple(dout.shape) != (B, Cn, Ho, Wo):
        raise ValueError(f"dout{tuple(dout.shape)} does not match the output shape {(B, Cn, Ho, Wo)}")
    if scale is None:
        scale = (D // heads) ** -0.5
    tap_tabs = device_tap_tables(Ho, Wo, h, w, K, dev)
    q = as_pixel_major(q.detach())
    k = _contig_pixel_major(k.detach())
    v = _contig_pixel_major(v.detach())
    dout = _contig_pixel_major(dout.detach())
    dq = torch.empty((B, Ho, Wo, D), device=dev, dtype=torch.float32)
    dk = torch.empty((B, h, w, D), device=dev, dtype=torch.float32)
    dv = torch.empty((B, h, w, Cn), device=dev, dtype=torch.float32)
    p = _lib.XAttnBwdParams()
    p.q, p.k, p.v, p.dout, p.dq, p.dk, p.dv = (_ptr(t) for t in (q, k, v, dout, dq, dk, dv))
    p.row_tap, p.col_tap = _ptr(tap_tabs[0]), _ptr(tap_tabs[1])
    if rope_tables is not None:
        p.cos_y, p.sin_y, p.cos_x, p.sin_x = (_ptr(t) for t in rope_tables)
    p.B, p.D, p.C, p.heads, p.Ho, p.Wo, p.h, p.w, p.K = B, D, Cn, heads, Ho, Wo, h, w, K[0]
    p.Kw = 0 if K[1] == K[0] else K[1]
    p.scale = float(scale)
    p.q_stride_b, _, p.q_stride_y, p.q_stride_x = q.stride()
    p.algo = int(algo)
    p.rep_y, p.rep_x = int(rep[0]), int(rep[1])
    with torch.cuda.device(dev):
        rc = _lib.load().naf_xattn_bwd_f32(C.byref(p), _stream(dev))
    _lib.check(rc, "naf_xattn_bwd_f32")
    return dq.permute(0, 3, 1, 2), dk.permute(0, 3, 1, 2), dv.permute(0, 3, 1, 2)


def rope_kpool_bwd(dq, dk, tables, rope_heads: int, inplace: bool = False):
    """Gradient of `rope_kpool` w.r.t. its input x (naf_rope_kpool_bwd_f32): dq (B,D,Ho,Wo) is the
    gradient w.r.t. the rotated map (or None), dk (B,D,h,w) w.r.t. the pooled keys (or None);
    tables as in the forward (None: no rotation).  Returns dx (B,D,Ho,Wo), pixel-major view."""
    dev = _require_cuda(dq, dk)
    if dq is None and dk is None:
        raise ValueError("rope_kpool_bwd: nothing to propagate")
    if dq is not None:
        dq = _contig_pixel_major(dq.detach())
        B, D, Ho, Wo = dq.shape
    if dk is not None:
        dk = _contig_pixel_major(dk.detach())
    if dq is None:
        if tables is None:
            raise ValueError("rope_kpool_bwd: the map size is unknown without dq or tables")
        B, D = dk.shape[:2]
        Ho, Wo = tables[0].shape[0], tables[2].shape[0]
    h, w = (dk.shape[-2], dk.shape[-1]) if dk is not None else (0, 0)
    dx = dq.permute(0, 2, 3, 1) if (inplace and dq is not None) else torch.empty((B, Ho, Wo, D), device=dev, dtype=torch.float32)
    p = _lib.KPoolBwdParams()
    p.dq, p.dk, p.dx = _ptr(dq), _ptr(dk), _ptr(dx)
    if tables is not None:
        p.cos_y, p.sin_y, p.cos_x, p.sin_x = (_ptr(t) for t in tables)
    p.B, p.D, p.Ho, p.Wo, p.h, p.w = B, D, Ho, Wo, h, w
    p.rope_heads = int(rope_heads)
    with torch.cuda.device(dev):
        rc = _lib.load().naf_rope_kpool_bwd_f32(C.byref(p), _stream(dev))
    _lib.check(rc, "naf_rope_kpool_bwd_f32")
    return dx.permute(0, 3, 1, 2)


def select_algo(q_shape, v_shape, heads: int, kernel_size, rope_on_the_fly: bool = True,
                return_scores: bool = False) -> str:
    """Name of the kernel AUTO would pick for these shapes (no launch; dummy aligned pointers)."""
    B, D, Ho, Wo = q_shape
    _, Cn, h, w = v_shape
    K = taps._pair(kernel_size)
    rt, _ = taps.tap_tables(Ho, Wo, h, w, K)
    p = _lib.XAttnParams()
    dummy = 1 << 20
    p.q = p.k = p.v = p.out = dummy
    if return_scores:
        p.scores = dummy
    if rt is not None:
        p.row_tap = p.col_tap = dummy
    if rope_on_the_fly:
        p.cos_y = p.sin_y = p.cos_x = p.sin_x = dummy
    p.B, p.D, p.C, p.heads, p.Ho, p.Wo, p.h, p.w, p.K = B, D, Cn, heads, Ho, Wo, h, w, K[0]
    p.Kw = 0 if K[1] == K[0] else K[1]
    p.scale = (D // heads) ** -0.5
    p.q_stride_b, p.q_stride_y, p.q_stride_x = Ho * Wo * D, Wo * D, D
    p.workspace, p.workspace_bytes = dummy, 1 << 62     # as ops.xattn provides it
    rc = _lib.load().naf_xattn_select_algo(C.byref(p))
    if rc < 0:
        _lib.check(-rc, "naf_xattn_select_algo")
    return _lib.ALGO_NAMES[rc]


def dump_taps(Ho: int, Wo: int, h: int, w: int, K: int, dev, force_tables: bool = False) -> torch.Tensor:
    """(Ho, Wo, K*K) int32 linear low-res indices the kernels gather from (test hook)."""
    if force_tables:
        rt = torch.from_numpy(taps.axis_taps(Ho, h, K).copy()).to(dev)
        ct = torch.from_numpy(taps.axis_taps(Wo, w, K).copy()).to(dev)
    else:
        rt, ct = device_tap_tables(Ho, Wo, h, w, K, dev)
    out = torch.empty((Ho, Wo, K * K), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        rc = _lib.load().naf_xattn_dump_taps_i32(_ptr(out), _ptr(rt), _ptr(ct), Ho, Wo, h, w, K, _stream(dev))
    _lib.check(rc, "naf_xattn_dump_taps_i32")
    return out
