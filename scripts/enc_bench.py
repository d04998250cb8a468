"""Time the tensor-core encoder kernels alone (CUDA events): python scripts/enc_bench.py [B] [H] [W] [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import encoder_fast

a = sys.argv[1:]
B, H, W, passes = (int(a[i]) if len(a) > i else d for i, d in ((0, 8), (1, 448), (2, 448), (3, 1)))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = naf_b200.NAF(kernel_size=7).eval().to(dev)
img = torch.randn(B, 3, H, W, device=dev)
x = torch.empty(B, H, W, 256, device=dev)
for name, seq, off in (("1x1 branch", m.image_encoder.encoder, 0), ("3x3 branch", m.image_encoder.sem_encoder, 128)):
    for _ in range(3):
        encoder_fast.forward_tc(seq, img, out=x, ch_off=off, passes=passes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        encoder_fast.forward_tc(seq, img, out=x, ch_off=off, passes=passes)
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: B={B} {H}x{W} passes={passes}: {e0.elapsed_time(e1) / n:.3f} ms per branch (stem + 4 fused GN/SiLU/conv layers)")
