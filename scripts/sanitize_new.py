"""Small launches of the round-2 kernels for compute-sanitizer (memcheck / racecheck): union-window forward
(tap tables, ratio 1, multi-chunk) and the tensor-core backward (1 and 3 value chunks, partial last tile)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from naf_b200 import _lib, ops


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32)).cuda()


for (B, D, n, C, Ho, Wo, h, w, K) in [(1, 256, 4, 64, 32, 32, 13, 13, 9), (1, 96, 1, 3, 24, 20, 24, 20, 15), (1, 64, 2, 10, 30, 45, 7, 11, 3)]:
    q, k, v = rnd(1, B, D, Ho, Wo), rnd(2, B, D, h, w), rnd(3, B, C, h, w)
    a = ops.xattn(q, k, v, n, K, algo=_lib.ALGO_UNION_TC)
    b = ops.xattn(q, k, v, n, K, algo=_lib.ALGO_GENERIC)
    print("union", (a - b).abs().max().item())
# backward: one-tile cells (1 and 3 value chunks), then multi-tile cells: 7 tiles with a second Q image / S region
# (r = 28, K = 3 and K = 7, dv = 192) and 2 tiles with one Q image (r = 16, K = 9: the reference's benchmark shape)
for (B, C, Ho, h, K) in [(1, 128, 40, 4, 3), (1, 768, 36, 4, 3), (1, 384, 72, 9, 9), (1, 768, 84, 3, 3), (1, 768, 196, 7, 7), (1, 384, 144, 9, 9)]:
    q, k, v, d = rnd(1, B, 256, Ho, Ho), rnd(2, B, 256, h, h), rnd(3, B, C, h, h), rnd(4, B, C, Ho, Ho)
    a = ops.xattn_bwd(q, k, v, d, 4, K, algo=_lib.ALGO_CELL_TC)
    b = ops.xattn_bwd(q, k, v, d, 4, K, algo=_lib.ALGO_GENERIC)
    print("bwd_tc", [(x - y).abs().max().item() for x, y in zip(a, b)])
torch.cuda.synchronize()
print("ok")
