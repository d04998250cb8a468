"""Autograd bindings of the CUDA forward / backward kernels.

The reference is differentiable through NATTEN's functionals (train.py:136, test/backward_speed.py:
51-64); these `torch.autograd.Function`s give `naf_b200.CrossAttention`, `naf_b200.RoPE` and `naf_b200.NAF`
the same property.  Forward = the inference kernels, bit for bit; backward = `naf_xattn_bwd_f32` and
`naf_rope_kpool_bwd_f32` (nothing but the inputs is saved: the probabilities are recomputed).
"""
from __future__ import annotations

import torch

from . import ops


class XAttnFn(torch.autograd.Function):
    """out = xattn(q, k, v): operator-level (CrossAttention.forward, src/layers/attentions.py:53-75)."""

    @staticmethod
    def forward(ctx, q, k, v, heads, kernel_size, scale, algo, out_dtype, return_scores):
        res = ops.xattn(q, k, v, heads, kernel_size, scale=scale, return_scores=return_scores, algo=algo,
                        out_dtype=out_dtype)
        ctx.save_for_backward(q, k, v)
        ctx.cfg = (heads, kernel_size, scale)
        if return_scores:
            ctx.mark_non_differentiable(res[1])
            return res
        return res

    @staticmethod
    def backward(ctx, dout, *unused):
        q, k, v = ctx.saved_tensors
        heads, kernel_size, scale = ctx.cfg
        dq, dk, dv = ops.xattn_bwd(q.float(), k.float(), v.float(), dout.float(), heads, kernel_size, scale=scale)
        return (dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None, None, None, None)


class RoPEFn(torch.autograd.Function):
    """q = RoPE(x) on its own (RoPE.forward, src/layers/rope.py:155-174)."""

    @staticmethod
    def forward(ctx, x, tables, rope_heads):
        _, q = ops.rope_kpool(x, tables, rope_heads, pooled_hw=None, want_q=True)
        ctx.tables, ctx.rope_heads = tables, rope_heads
        return q

    @staticmethod
    def backward(ctx, dq):
        return ops.rope_kpool_bwd(dq.float(), None, ctx.tables, ctx.rope_heads), None, None


class KeyPoolFn(torch.autograd.Function):
    """k = adaptive_avg_pool2d(q, (h, w)) (KeyEncoder.forward, src/model/naf.py:63-69)."""

    @staticmethod
    def forward(ctx, q, hw):
        k, _ = ops.rope_kpool(q, None, 1, pooled_hw=hw, want_q=False)
        ctx.shape = tuple(q.shape)
        return k

    @staticmethod
    def backward(ctx, dk):
        B, D, Ho, Wo = ctx.shape
        zero = torch.zeros((B, Ho, Wo, D), device=dk.device, dtype=torch.float32).permute(0, 3, 1, 2)
        return ops.rope_kpool_bwd(zero, dk.float(), None, 1, inplace=True), None


class NAFUpsampleFn(torch.autograd.Function):
    """The fused hot path of NAF.forward downstream of the conv encoder (src/model/naf.py:104-116):
    un-rotated pooled guidance x + features -> out, RoPE applied on the fly, keys pooled from the
    rotated map.  Backward: dV scatter-add, dQ through the transposed rotation, dK through the pooling."""

    @staticmethod
    def forward(ctx, x, features, tables, pool_tables, rope_heads, heads, kernel_size, scale, algo, rep, out_dtype):
        x = ops.as_pixel_major(x)
        h, w = features.shape[-2:]
        if pool_tables is not None:      # replicated guidance: pool the source map with block-mean tables
            k, _ = ops.rope_kpool(x, pool_tables, rope_heads, pooled_hw=(h, w), want_q=False)
        else:
            k, _ = ops.rope_kpool(x, tables, rope_heads, pooled_hw=(h, w), want_q=False, rep=rep)
        out = ops.xattn(x, k, features, heads, kernel_size, scale=scale, rope_tables=tables, algo=algo, rep=rep,
                        out_dtype=out_dtype)
        ctx.save_for_backward(x, k, features)
        ctx.cfg = (tables, rope_heads, heads, kernel_size, scale, rep)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, k, features = ctx.saved_tensors
        tables, rope_heads, heads, kernel_size, scale, rep = ctx.cfg
        dq, dk, dv = ops.xattn_bwd(x, k, features.float(), dout.float(), heads, kernel_size, scale=scale,
                                   rope_tables=tables, rep=rep)
        dx = ops.rope_kpool_bwd(dq, dk, tables, rope_heads, inplace=True)      # at the target resolution
        ry, rx = int(rep[0]), int(rep[1])
        if ry > 1 or rx > 1:
            # x was a replicated source map: every source pixel collects its ry x rx block (the backward of
            # the reference's adaptive_avg_pool2d up-replication, src/model/naf.py:34)
            B, D, Ho, Wo = dx.shape
            dx = dx.reshape(B, D, Ho // ry, ry, Wo // rx, rx).sum(dim=(3, 5))
        return (dx, dv.to(features.dtype), None, None, None, None, None, None, None, None, None)
