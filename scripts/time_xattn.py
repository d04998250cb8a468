"""Time the attention kernel alone (C2 shape by default) with CUDA events.
   python scripts/time_xattn.py [B] [algo] [C] [to] [lo] [K] [rep]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import ops

a = sys.argv[1:]
B = int(a[0]) if len(a) > 0 else 8
algo = {"auto": 0, "generic": 1, "cell_simt": 2, "cell_tc": 3, "cell_tcws": 4, "cell_tma": 5}[a[1] if len(a) > 1 else "auto"]
C, to, lo, K, rep = (int(a[i]) if len(a) > i else d for i, d in ((2, 768), (3, 896), (4, 32), (5, 7), (6, 1)))
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        out = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=algo, rep=(rep, rep))
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
gb = 4.0 * B * (256 * to * to + C * lo * lo + C * to * to) / 1e9
print(f"lib={os.environ.get('NAF_B200_LIB','default')} B={B} C={C} {to}/{lo} K={K} rep={rep}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s algorithmic  ({gb/ms*1e3/6534.1*100:.1f}% of 6534)")
