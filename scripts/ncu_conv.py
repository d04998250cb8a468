"""Driver for ncu captures of the encoder kernels: both conv branches, twice (C2 guidance size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import encoder_fast

B, H, W = 8, 448, 448
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = naf_b200.NAF(kernel_size=7).eval().to(dev)
img = torch.randn(B, 3, H, W, device=dev)
x = torch.empty(B, H, W, 256, device=dev)
for _ in range(2):
    encoder_fast.forward_tc(m.image_encoder.encoder, img, out=x, ch_off=0, passes=passes)
    encoder_fast.forward_tc(m.image_encoder.sem_encoder, img, out=x, ch_off=128, passes=passes)
torch.cuda.synchronize()
print("ok")
