"""One backward launch of the tensor-core cell kernel on a given shape (compute-sanitizer target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from naf_b200 import _lib, ops

B, C, Ho, h, K = (int(a) for a in sys.argv[1:6])
rnd = lambda seed, *shape: torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32)).cuda()
q, k, v, d = rnd(1, B, 256, Ho, Ho), rnd(2, B, 256, h, h), rnd(3, B, C, h, h), rnd(4, B, C, Ho, Ho)
a = ops.xattn_bwd(q, k, v, d, 4, K, algo=_lib.ALGO_CELL_TC)
torch.cuda.synchronize()
b = ops.xattn_bwd(q, k, v, d, 4, K, algo=_lib.ALGO_GENERIC)
print("bwd_tc vs generic", [(x - y).abs().max().item() for x, y in zip(a, b)])
