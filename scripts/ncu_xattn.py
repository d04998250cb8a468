"""Single-kernel driver for ncu captures: one BASELINE-shaped attention launch, nothing else.
   python scripts/ncu_xattn.py B algo rep [C to lo K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import _lib, ops

a = sys.argv[1:]
B = int(a[0]) if len(a) > 0 else 2
algo = {"auto": 0, "generic": 1, "cell_simt": 2, "cell_tcws": 4, "cell_tma": 5}[a[1] if len(a) > 1 else "auto"]
rep = int(a[2]) if len(a) > 2 else 1
C, to, lo, K = (int(a[i]) if len(a) > i else d for i, d in ((3, 768), (4, 896), (5, 32), (6, 7)))
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
