"""Driver for ncu captures of the backward attention kernel: C2-like cells (r = 28, K = 7, dv = 192), one image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import _lib, ops

B, D, n, C, Ho, h, K = 1, 256, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 768, 896, 32, 7
dev = torch.device("cuda", 0)
torch.manual_seed(0)
q = torch.randn(B, D, Ho, Ho, device=dev)
k = torch.randn(B, D, h, h, device=dev)
v = torch.randn(B, C, h, h, device=dev)
dout = torch.randn(B, C, Ho, Ho, device=dev)
tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(Ho, Ho)
for _ in range(2):
    ops.xattn_bwd(q, k, v, dout, n, K, rope_tables=tabs, algo=_lib.ALGO_CELL_TC)
torch.cuda.synchronize()
print("ok")
