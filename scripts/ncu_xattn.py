"""Single-kernel driver for ncu captures: C2-shaped attention launch (B images), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
algo = {"auto": 0, "generic": 1, "cell_simt": 2, "cell_tc": 3, "cell_tcws": 4, "cell_tma": 5}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
C, to, lo, K = 768, 896, 32, 7
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(B, to // rep, to // rep, 256, device=dev).permute(0, 3, 1, 2)
feats = torch.randn(B, C, lo, lo, device=dev)
m = naf_b200.NAF(kernel_size=K).eval().to(dev)
tables = m.image_encoder.rope.axis_tables(to, to)
with torch.no_grad():
    k, _ = ops.rope_kpool(x, tables, 4, pooled_hw=(lo, lo), rep=(rep, rep))
    for _ in range(3):
        out = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=algo, rep=(rep, rep))
torch.cuda.synchronize()
print("ok", out.shape)
