"""Driver for an ncu capture of the union-window kernel: the denoising shape (ratio 1, one head of 256, C = 3, K = 15)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, D, n, C, S, K = 2, 256, 1, 3, 256, 15
q, k, v = torch.randn(B, D, S, S, device=dev), torch.randn(B, D, S, S, device=dev), torch.randn(B, C, S, S, device=dev)
tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(S, S)
for _ in range(3):
    ops.xattn(q, k, v, n, K, rope_tables=tabs)
torch.cuda.synchronize()
print("ok")
