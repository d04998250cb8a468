"""Axial 2-D rotary position embedding (DINOv3 style), CUDA-backed.

Mirrors the reference operator `RoPE` (src/layers/rope.py:39-174): same constructor keywords, the
same persistent `periods` buffer (so `state_dict` keys match), the same `forward(x, layout)`.
The rotation itself runs in `naf_rope_kpool_f32`; inside `NAF.forward` it is fused further (the
attention kernel rotates q on the fly and the key pre-pass rotates while it pools), so this
module's `forward` is only used when somebody calls RoPE on its own.

Coordinates follow `create_coordinate` (src/layers/rope.py:84-126) operation for operation: the three
normalisations ("separate", "min", "max") and the train-time augmentations (`shift_coords`,
`jitter_coords`, `rescale_coords`: random draws made with the same torch calls, in the same order, on
the same device as the reference, so a seeded run draws the same numbers).  Every one of them acts per
axis, so the coordinates -- and the cos/sin tables -- stay factored into a row part and a column part.
Like the reference (:159-161) the coordinates are cached per (H, W): they are re-drawn only when the map
size changes, whatever the training flag says.
"""
from __future__ import annotations

import math
from typing import Literal, Optional

import torch
from torch import Tensor, nn

from .. import ops


class RoPE(nn.Module):
    def __init__(
        self,
        embed_dim: int,
        *,
        num_heads: int,
        base: Optional[float] = 100.0,
        min_period: Optional[float] = None,
        max_period: Optional[float] = None,
        normalize_coords: Literal["min", "max", "separate"] = "separate",
        shift_coords: Optional[float] = None,
        jitter_coords: Optional[float] = None,
        rescale_coords: Optional[float] = None,
        dtype: Optional[torch.dtype] = None,
        device: Optional[torch.device] = None,
    ):
        super().__init__()
        assert embed_dim % (4 * num_heads) == 0
        has_range = min_period is not None and max_period is not None
        if (base is None) == (not has_range):
            raise ValueError("Either `base` or `min_period`+`max_period` must be provided.")
        self.num_heads = num_heads
        self.base = base
        self.min_period = min_period
        self.max_period = max_period
        self.D_head = embed_dim // num_heads
        self.normalize_coords = normalize_coords
        self.shift_coords = shift_coords
        self.jitter_coords = jitter_coords
        self.rescale_coords = rescale_coords
        self.dtype = dtype
        self.register_buffer("periods", torch.empty(self.D_head // 4, device=device, dtype=dtype),
                             persistent=True)
        self._init_weights()
        self._table_key = None
        self._tables = None

    def _init_weights(self) -> None:
        n = self.D_head // 4
        dev = self.periods.device
        if self.base is not None:
            exponent = 2 * torch.arange(n, device=dev, dtype=self.dtype) / (self.D_head // 2)
            periods = self.base ** exponent
        else:
            periods = torch.logspace(math.log10(self.min_period), math.log10(self.max_period), steps=n)
        self.periods.data = periods

    # -- coordinates and tables --------------------------------------------------------------
    def axis_coords(self, H: int, W: int):
        """(coords_y (H,), coords_x (W,)) in [-1, 1]: `create_coordinate` of the reference, per axis."""
        dev, dt = self.periods.device, self.dtype
        dd = {"device": dev, "dtype": dt}
        if self.normalize_coords == "max":
            den_h = den_w = max(H, W)
        elif self.normalize_coords == "min":
            den_h = den_w = min(H, W)
        elif self.normalize_coords == "separate":
            den_h, den_w = H, W
        else:
            raise ValueError(f"Unknown normalize_coords: {self.normalize_coords}")
        cy = 2.0 * (torch.arange(0.5, H, **dd) / den_h) - 1.0
        cx = 2.0 * (torch.arange(0.5, W, **dd) / den_w) - 1.0
        if self.training and self.shift_coords is not None:
            shift_hw = torch.empty(2, **dd).uniform_(-self.shift_coords, self.shift_coords)
            cy, cx = cy + shift_hw[0], cx + shift_hw[1]
        if self.training and self.jitter_coords is not None:
            jitter_max = math.log(self.jitter_coords)
            jitter_hw = torch.empty(2, **dd).uniform_(-jitter_max, jitter_max).exp()
            cy, cx = cy * jitter_hw[0], cx * jitter_hw[1]
        if self.training and self.rescale_coords is not None:
            rescale_max = math.log(self.rescale_coords)
            rescale_hw = torch.empty(1, **dd).uniform_(-rescale_max, rescale_max).exp()
            cy, cx = cy * rescale_hw, cx * rescale_hw
        return cy, cx

    def axis_tables(self, H: int, W: int):
        """(cos_y, sin_y, cos_x, sin_x) for an (H, W) map, cached per shape / device / periods (the
        reference caches its coordinates per (H, W) too, src/layers/rope.py:159-161)."""
        key = (H, W, self.periods.device, self.periods.data_ptr(), self.periods._version)
        if key != self._table_key:
            cy, cx = self.axis_coords(H, W)
            self._tables = ops.rope_tables_from_coords(cy.to(torch.float32), cx.to(torch.float32),
                                                       self.periods.to(torch.float32))
            self._table_key = key
        return self._tables

    def mean_axis_tables(self, H: int, W: int, ry: int, rx: int):
        """Tables at (H/ry, W/rx) holding the MEAN cos/sin over each block of ry rows / rx columns
        (used to pool keys directly from a replicated guidance map; see NAF.upsample_from_guidance)."""
        cy, sy, cx, sx = self.axis_tables(H, W)
        key = (self._table_key, ry, rx)
        if getattr(self, "_mean_key", None) != key:
            P = cy.shape[1]
            self._mean_tables = (cy.view(H // ry, ry, P).mean(1).contiguous(), sy.view(H // ry, ry, P).mean(1).contiguous(),
                                 cx.view(W // rx, rx, P).mean(1).contiguous(), sx.view(W // rx, rx, P).mean(1).contiguous())
            self._mean_key = key
        return self._mean_tables

    def forward(self, x: Tensor, layout: str = "spatial") -> Tensor:
        B, D, H, W = x.shape
        if D != self.D_head * self.num_heads:
            raise ValueError(f"expected {self.D_head * self.num_heads} channels, got {D}")
        tables = self.axis_tables(H, W)
        if torch.is_grad_enabled() and x.requires_grad:
            from ..autograd import RoPEFn
            q = RoPEFn.apply(x, tables, self.num_heads)
        else:
            _, q = ops.rope_kpool(x, tables, self.num_heads, pooled_hw=None, want_q=True)
        if layout == "spatial":
            return q
        if layout == "flatten":
            return q.permute(0, 2, 3, 1).reshape(B, H, W, self.num_heads, self.D_head)
        return q.permute(0, 2, 3, 1).reshape(B, H * W, self.num_heads, self.D_head).permute(0, 2, 1, 3)
