"""TEST INFRASTRUCTURE ONLY -- a pure-torch stand-in for the two NATTEN functionals the
reference calls, so that the UNMODIFIED reference modules can be imported and run on CPU.

The reference's arithmetic core is the third-party library NATTEN (PyPI ``natten``, pinned
``0.17.4+torch240cu118`` for development, also ``0.17.3`` / ``0.20.1``; see
/root/reference/docs/INSTALL.md:7,16,20,35 and /root/reference/hubconf.py:1).  It is not under
/root/reference, not installed in this image and not installable offline.  This module restates
NATTEN 0.17.x's *published* 2-D neighbourhood-attention semantics for the two call sites

    /root/reference/src/layers/attentions.py:20   na2d_qk(q, k, kernel_size=, dilation=)
    /root/reference/src/layers/attentions.py:24   na2d_av(attn, v, kernel_size=, dilation=)

on ``(B, heads, H, W, d)`` tensors:

* the window of a pixel is ``kernel_size`` taps per axis, spaced ``dilation`` apart, *shifted*
  (never truncated or padded) so that it stays inside the map, within the pixel's dilation group
  ``i % dilation`` (NATTEN's ``get_window_start``);
* tap order of the attention tensor is row-major ``t_h * K_w + t_w``;
* NATTEN raises when ``kernel_size * dilation`` exceeds the axis length.

PARITY STATUS: the window rule is restated from the library's documentation / 0.17 source as
recalled in SURVEY.md (no NATTEN binary exists here to pin it against): "parity unpinned" at
this one boundary.  It is kept in ONE function (`window_start`) so it can be swapped if a real
NATTEN ever becomes available.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this file.  Nothing under ``naf_b200/`` does.
"""
from __future__ import annotations

import sys
import types

import torch


def window_start(i: int, length: int, kernel: int, dilation: int) -> int:
    """First tap of pixel ``i`` along an axis of ``length`` (NATTEN 0.17 `get_window_start`)."""
    ns = kernel // 2
    d = dilation
    if d <= 1:
        return max(i - ns, 0) + ((length - i - ns - 1) if (i + ns >= length) else 0)
    ni = i - ns * d
    if ni < 0:
        return i % d
    if i + ns * d >= length:
        imodd = i % d
        a = (length // d) * d
        b = length - a
        if imodd < b:
            return length - b + imodd - 2 * ns * d
        return a + imodd - kernel * d
    return ni


def axis_taps(length: int, kernel: int, dilation: int) -> torch.Tensor:
    """(length, kernel) int64 table: tap positions of every pixel along one axis."""
    if kernel * dilation > length:
        raise ValueError(
            f"natten stub: kernel_size*dilation ({kernel}*{dilation}) exceeds axis length {length}"
        )
    if kernel % 2 != 1 or kernel < 1:
        raise ValueError(f"natten stub: kernel_size must be odd and positive, got {kernel}")
    rows = []
    for i in range(length):
        s = window_start(i, length, kernel, dilation)
        rows.append([s + t * dilation for t in range(kernel)])
    return torch.tensor(rows, dtype=torch.int64)


def _pair(v):
    if isinstance(v, (tuple, list)):
        assert len(v) == 2
        return int(v[0]), int(v[1])
    return int(v), int(v)


_ROWS_PER_CHUNK = 16


def na2d_qk(q: torch.Tensor, k: torch.Tensor, kernel_size, dilation=1) -> torch.Tensor:
    """(B,n,H,W,d),(B,n,H,W,d) -> (B,n,H,W,Kh*Kw) raw (unscaled) neighbourhood logits."""
    kh, kw = _pair(kernel_size)
    dh, dw = _pair(dilation)
    B, n, H, W, d = q.shape
    assert k.shape == q.shape, (k.shape, q.shape)
    rt = axis_taps(H, kh, dh).to(q.device)  # (H,kh)
    ct = axis_taps(W, kw, dw).to(q.device)  # (W,kw)
    out = q.new_empty(B, n, H, W, kh * kw)
    for y0 in range(0, H, _ROWS_PER_CHUNK):
        y1 = min(H, y0 + _ROWS_PER_CHUNK)
        r = rt[y0:y1]  # (hc,kh)
        # neigh: (B,n,hc,kh,W,kw,d)
        neigh = k[:, :, r[:, :, None, None], ct[None, None, :, :], :]
        qq = q[:, :, y0:y1, None, :, None, :]  # (B,n,hc,1,W,1,d)
        s = (qq * neigh).sum(-1)  # (B,n,hc,kh,W,kw)
        out[:, :, y0:y1] = s.permute(0, 1, 2, 4, 3, 5).reshape(B, n, y1 - y0, W, kh * kw)
    return out


def na2d_av(attn: torch.Tensor, v: torch.Tensor, kernel_size, dilation=1) -> torch.Tensor:
    """(B,n,H,W,Kh*Kw),(B,n,H,W,dv) -> (B,n,H,W,dv) weighted neighbourhood sum."""
    kh, kw = _pair(kernel_size)
    dh, dw = _pair(dilation)
    B, n, H, W, dv = v.shape
    assert attn.shape == (B, n, H, W, kh * kw), (attn.shape, v.shape)
    rt = axis_taps(H, kh, dh).to(v.device)
    ct = axis_taps(W, kw, dw).to(v.device)
    out = v.new_empty(B, n, H, W, dv)
    for y0 in range(0, H, _ROWS_PER_CHUNK):
        y1 = min(H, y0 + _ROWS_PER_CHUNK)
        r = rt[y0:y1]
        neigh = v[:, :, r[:, :, None, None], ct[None, None, :, :], :]  # (B,n,hc,kh,W,kw,dv)
        a = attn[:, :, y0:y1].reshape(B, n, y1 - y0, W, kh, kw).permute(0, 1, 2, 4, 3, 5)
        out[:, :, y0:y1] = (a[..., None] * neigh).sum(dim=(3, 5))
    return out


def install() -> None:
    """Register `natten` / `natten.functional` stand-ins in sys.modules (idempotent)."""
    if "natten" in sys.modules and getattr(sys.modules["natten"], "__naf_stub__", False):
        return
    pkg = types.ModuleType("natten")
    pkg.__naf_stub__ = True
    pkg.__version__ = "0.17.4+stub"
    fn = types.ModuleType("natten.functional")
    fn.na2d_qk = na2d_qk
    fn.na2d_av = na2d_av
    pkg.functional = fn
    sys.modules["natten"] = pkg
    sys.modules["natten.functional"] = fn
