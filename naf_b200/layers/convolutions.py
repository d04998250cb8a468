"""Guidance-image conv encoder (the feeder of the hot path).

Same module tree -- and therefore the same `state_dict` keys and shapes -- as the reference's
`encoder()` / `EncBlock` (src/layers/convolutions.py:9-95): a stem conv followed by pre-norm
blocks GroupNorm -> SiLU -> conv -> GroupNorm -> SiLU -> conv with reflect padding.  The modules
hold the parameters (and run as plain torch modules on CPU tensors, in training mode and for
non-default widths); on CUDA `NAF.forward` evaluates the reference-default stacks through the fused
tcgen05 kernels of `naf_b200/encoder_fast.py:forward_tc` instead of calling these modules.
"""
from __future__ import annotations

from torch import nn


def _conv(cin, cout, k, pad_mode, bias):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, padding_mode=pad_mode, bias=bias)


class EncBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, norm_kwargs={}, pad_mode="zeros",
                 norm_fn=None, activation_fn=nn.SiLU, use_conv_shortcut=False, bias=True,
                 residual=False):
        super().__init__()
        self.use_conv_shortcut = use_conv_shortcut
        self.residual = residual
        self.norm1 = norm_fn(**norm_kwargs)
        self.conv1 = _conv(in_channels, out_channels, kernel_size, pad_mode, bias)
        self.norm2 = norm_fn(**norm_kwargs)
        self.conv2 = _conv(out_channels, out_channels, kernel_size, pad_mode, bias)
        self.activation_fn = activation_fn()
        if in_channels != out_channels:
            self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=bias)

    def forward(self, x):
        y = self.conv1(self.activation_fn(self.norm1(x)))
        y = self.conv2(self.activation_fn(self.norm2(y)))
        if not self.residual:
            return y
        skip = x
        if self.use_conv_shortcut or skip.shape != y.shape:
            skip = self.shortcut(skip)
        return y + skip


def encoder(in_dim, hidden_dim, kernel_size=1, ks_res=1, num_layers=2, bias=True, num_groups=8,
            residual=False):
    blocks = [
        EncBlock(hidden_dim, hidden_dim, kernel_size=ks_res, pad_mode="reflect",
                 norm_fn=nn.GroupNorm,
                 norm_kwargs={"num_groups": num_groups, "num_channels": hidden_dim},
                 activation_fn=nn.SiLU, use_conv_shortcut=False, bias=bias, residual=residual)
        for _ in range(num_layers)
    ]
    return nn.Sequential(_conv(in_dim, hidden_dim, kernel_size, "reflect", bias), *blocks)
