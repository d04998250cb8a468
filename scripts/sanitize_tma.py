"""Small launches of the TMA attention kernel for compute-sanitizer (memcheck / racecheck): RoPE table rows staged in
shared memory (bulk copies + chunk-plane tensor loads), replicated guidance, several items per CTA (more than two, so
every buffer is re-used), partial last tile, K = 11 / dv = 256 (tables from global), return_weights; and the generic
kernels with a rectangular window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import naf_b200
from naf_b200 import _lib, ops


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32)).cuda()


m = naf_b200.NAF(kernel_size=7).eval().cuda()
# (B, C, Ho, h, K, rep): 2*20*20*4 = 3200 items > 2 * 148 CTAs; 14-pixel cells: 2 tiles, the last one partial
for (B, C, Ho, h, K, rep) in [(2, 768, 280, 20, 7, 2), (1, 384, 196, 14, 5, 1), (1, 1024, 308, 11, 11, 2), (1, 128, 180, 18, 3, 1)]:
    with torch.no_grad():
        x = rnd(1, B, 256, Ho // rep, Ho // rep)
        feats = rnd(2, B, C, h, h)
        tables = m.image_encoder.rope.axis_tables(Ho, Ho)
        k, _ = ops.rope_kpool(x, tables, 4, pooled_hw=(h, h), rep=(rep, rep))
        a = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=_lib.ALGO_CELL_TMA, rep=(rep, rep))
        b = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=_lib.ALGO_CELL_SIMT, rep=(rep, rep))
        print("tma", (B, C, Ho, h, K, rep), (a - b).abs().max().item())
with torch.no_grad():
    x, feats = rnd(3, 1, 256, 112, 112), rnd(4, 1, 256, 8, 8)
    tables = m.image_encoder.rope.axis_tables(112, 112)
    k, _ = ops.rope_kpool(x, tables, 4, pooled_hw=(8, 8))
    a, sa = ops.xattn(x, k, feats, 4, 7, rope_tables=tables, algo=_lib.ALGO_CELL_TMA, return_scores=True)
    b, sb = ops.xattn(x, k, feats, 4, 7, rope_tables=tables, algo=_lib.ALGO_GENERIC, return_scores=True)
    print("tma scores", (a - b).abs().max().item(), (sa - sb).abs().max().item())
    q, kk, v, d = rnd(5, 1, 64, 30, 45), rnd(6, 1, 64, 7, 11), rnd(7, 1, 10, 7, 11), rnd(8, 1, 10, 30, 45)
    o = ops.xattn(q, kk, v, 2, (5, 3))
    g = ops.xattn_bwd(q, kk, v, d, 2, (5, 3))
    print("rect", o.abs().max().item(), [t.abs().max().item() for t in g])
torch.cuda.synchronize()
print("ok")
