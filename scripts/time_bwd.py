"""Time the backward attention kernel (tensor-core cell kernel) on two shapes; used with NAF_B200_LIB variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import _lib, ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (B, C, Ho, h, K) in [(1, 128, 448, 28, 9), (1, 384, 448, 28, 9), (1, 1024, 448, 28, 9), (1, 768, 896, 32, 7)]:
    D, n = 256, 4
    # pixel-major storage (what the autograd functions hand over): no packing pass inside the timed call
    q = torch.randn(B, Ho, Ho, D, device=dev).permute(0, 3, 1, 2); k = torch.randn(B, h, h, D, device=dev).permute(0, 3, 1, 2)
    v = torch.randn(B, h, h, C, device=dev).permute(0, 3, 1, 2); dout = torch.randn(B, Ho, Ho, C, device=dev).permute(0, 3, 1, 2)
    tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(Ho, Ho)
    fn = lambda: ops.xattn_bwd(q, k, v, dout, n, K, rope_tables=tabs, algo=_lib.ALGO_CELL_TC)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('NAF_B200_LIB', 'regular'):40s} C={C:5d} {Ho}<-{h} K={K}: {e0.elapsed_time(e1) / 5:8.3f} ms", flush=True)
