"""Cross-scale neighbourhood attention operator, CUDA-backed.

Drop-in for the reference `CrossAttention` (src/layers/attentions.py:32-75): same constructor,
same `forward(q, k, v, image=None, return_weights=False)`, same side effect (`self.dilation`),
same output -- a (B, C, Ho, Wo) permuted view over pixel-major storage.  Where the reference
replicates K and V to the target resolution (`_resize`, :48-51) and calls NATTEN on the dilated
high-resolution grid (:20,24,72), this operator hands the low-resolution K and V straight to
`naf_xattn_fwd_f32`, which evaluates the identical neighbourhoods through integer tap tables
(`naf_b200/taps.py`).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib, ops


def autocast_out_dtype(device_type: str = "cuda") -> torch.dtype:
    """Output dtype the reference would produce here: bf16 inside `torch.autocast(bfloat16)` (its
    NATTEN calls then run in bf16: train.py:120, denoising.py:209), fp32 otherwise.  fp16 autocast
    is not mirrored (the kernels offer fp32 and bf16 stores)."""
    if torch.is_autocast_enabled(device_type) and torch.get_autocast_dtype(device_type) == torch.bfloat16:
        return torch.bfloat16
    return torch.float32


class CrossAttention(nn.Module):
    def __init__(self, dim, num_heads, kernel_size=(9, 9), **kwargs):
        super().__init__()
        assert dim % num_heads == 0, "dim must be divisible by num_heads"
        self.num_heads = num_heads
        self.kernel_size = kernel_size
        self.scale = (dim // num_heads) ** -0.5
        self.algo = _lib.ALGO_AUTO

    def _window(self):
        """kernel_size as the kernels take it: an int for square windows, NATTEN's (kh, kw) pair otherwise
        (rectangular windows run on the generic kernel)."""
        ks = self.kernel_size
        if isinstance(ks, (tuple, list)):
            if len(ks) != 2:
                raise ValueError(f"kernel_size must be an int or a pair, got {ks!r}")
            kh, kw = int(ks[0]), int(ks[1])
            return kh if kh == kw else (kh, kw)
        return int(ks)

    def forward(self, q, k, v, image=None, return_weights=False, rope_tables=None, rep=(1, 1), out_dtype=None,
                **kwargs):
        hq, wq = q.shape[-2] * int(rep[0]), q.shape[-1] * int(rep[1])
        hk, wk = k.shape[-2:]
        self.dilation = (hq // hk, wq // wk)
        out_dtype = autocast_out_dtype() if out_dtype is None else out_dtype
        if (torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v)) and rope_tables is None
                and tuple(rep) == (1, 1)):
            # differentiable like the reference's NATTEN calls (train.py:136): forward = the same kernels
            from ..autograd import XAttnFn
            return XAttnFn.apply(q, k, v, self.num_heads, self._window(), self.scale, self.algo,
                                 out_dtype, bool(return_weights))
        res = ops.xattn(q, k, v, self.num_heads, self._window(), scale=self.scale,
                        rope_tables=rope_tables, return_scores=return_weights, algo=self.algo, rep=rep,
                        out_dtype=out_dtype)
        return res
