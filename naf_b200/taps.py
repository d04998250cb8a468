"""Host-side neighbourhood index tables.

For target index ``i`` along one axis the reference gathers ``K`` low-resolution indices: NATTEN's
dilated window (``dilation = Ho // h``, reference src/layers/attentions.py:54-56) slides over the
*target* grid, shifted so it never leaves the map, and every tap lands on the low-res cell that
``F.interpolate(mode="nearest-exact")`` replicated there (src/layers/attentions.py:48-51).  The
composition is a pure integer function of ``(Ho, h, K)``; it is evaluated here once per shape
(vectorised numpy), cached, and shipped to the device as an ``int32 (Ho, K)`` table.

When ``Ho`` is a multiple of ``h`` the table is block constant --
``clamp(i // r - K // 2, 0, h - K) + t`` -- and the kernels use that closed form instead
(`integer_ratio`), which is what lets a whole low-res cell share one shared-memory window.
"""
from __future__ import annotations

import functools

import numpy as np


def _pair(k) -> tuple[int, int]:
    if isinstance(k, (tuple, list)):
        if len(k) != 2:
            raise ValueError(f"kernel_size must be an int or a pair, got {k!r}")
        return int(k[0]), int(k[1])
    return int(k), int(k)


def nearest_exact_source(out_len: int, in_len: int) -> np.ndarray:
    """Index map of ATen's upsample_nearest_exact: min(floor((dst + 0.5) * scale), in - 1),
    scale = float32(in) / out, all in fp32 like the ATen kernel."""
    scale = np.float32(in_len) / np.float32(out_len)
    centres = np.arange(out_len, dtype=np.float32) + np.float32(0.5)
    return np.minimum(np.floor(centres * scale).astype(np.int64), in_len - 1)


def window_starts(length: int, kernel: int, dilation: int) -> np.ndarray:
    """First tap of every position along an axis (NATTEN shifted, dilation-grouped window)."""
    i = np.arange(length, dtype=np.int64)
    half = kernel // 2
    if dilation <= 1:
        start = np.maximum(i - half, 0)
        over = i + half >= length
        return start + np.where(over, length - i - half - 1, 0)
    reach = half * dilation
    group = i % dilation
    full = (length // dilation) * dilation
    rem = length - full
    right = np.where(group < rem, length - rem + group - 2 * reach, full + group - kernel * dilation)
    return np.where(i - reach < 0, group, np.where(i + reach >= length, right, i - reach))


def validate_window(out_len: int, in_len: int, kernel: int) -> int:
    """Checks NATTEN's constraints; returns the dilation the reference would use."""
    if kernel < 1 or kernel % 2 == 0:
        raise ValueError(f"kernel_size must be odd and >= 1, got {kernel}")
    if in_len < 1 or out_len < in_len:
        raise ValueError(f"target length {out_len} must be >= feature length {in_len} >= 1")
    dilation = out_len // in_len
    if kernel * dilation > out_len:
        raise ValueError(
            f"kernel_size*dilation exceeds the target size ({kernel}*{dilation} > {out_len})")
    return dilation


@functools.lru_cache(maxsize=256)
def axis_taps(out_len: int, in_len: int, kernel: int) -> np.ndarray:
    """(out_len, kernel) int32 low-res index of every tap of every target position."""
    dilation = validate_window(out_len, in_len, kernel)
    start = window_starts(out_len, kernel, dilation)
    hi_res = start[:, None] + np.arange(kernel, dtype=np.int64)[None, :] * dilation
    table = nearest_exact_source(out_len, in_len)[hi_res].astype(np.int32)
    table.setflags(write=False)
    return table


def integer_ratio(out_len: int, in_len: int) -> bool:
    return out_len % in_len == 0


def closed_form_axis_taps(out_len: int, in_len: int, kernel: int) -> np.ndarray:
    """The integer-ratio rule the cell kernels evaluate on the device."""
    assert integer_ratio(out_len, in_len)
    r = out_len // in_len
    cell = np.arange(out_len, dtype=np.int64) // r
    origin = np.clip(cell - kernel // 2, 0, in_len - kernel)
    return (origin[:, None] + np.arange(kernel, dtype=np.int64)[None, :]).astype(np.int32)


def tap_tables(Ho: int, Wo: int, h: int, w: int, kernel_size):
    """(row_tap, col_tap) numpy int32 tables, or (None, None) when both axes have an integer
    ratio (the device closed form is then exact; see tests/test_taps.py)."""
    kh, kw = _pair(kernel_size)
    validate_window(Ho, h, kh)
    validate_window(Wo, w, kw)
    if integer_ratio(Ho, h) and integer_ratio(Wo, w):
        return None, None
    return axis_taps(Ho, h, kh), axis_taps(Wo, w, kw)
