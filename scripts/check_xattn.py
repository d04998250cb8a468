"""Quick GPU correctness check of one attention algo against the SIMT kernel (no oracle needed).
   python scripts/check_xattn.py algo [B C to lo K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import ops
a = sys.argv[1:]
algo = {"auto": 0, "generic": 1, "cell_simt": 2, "cell_tc": 3, "cell_tcws": 4, "cell_tma": 5}[a[0]]
B, C, to, lo, K = (int(a[i]) if len(a) > i else d for i, d in ((1, 2), (2, 768), (3, 224), (4, 8), (5, 7)))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = (torch.randn(B, to, to, 256, device=dev) * 2).permute(0, 3, 1, 2)
feats = torch.randn(B, C, lo, lo, device=dev)
m = naf_b200.NAF(kernel_size=K).eval().to(dev)
tables = m.image_encoder.rope.axis_tables(to, to)
with torch.no_grad():
    k, _ = ops.rope_kpool(x, tables, 4, pooled_hw=(lo, lo))
    ref = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=2)
    out = ops.xattn(x, k, feats, 4, K, rope_tables=tables, algo=algo)
    torch.cuda.synchronize()
err = (out - ref).abs()
print(f"algo={a[0]} B={B} C={C} {to}/{lo} K={K}: max|err|={err.max().item():.3e} mean={err.mean().item():.3e} nan={torch.isnan(out).sum().item()}")
if err.max().item() > 1e-3:
    r = to // lo
    e = err.reshape(B, 4, C // 4, lo, r, lo, r).amax(dim=(2, 4, 6))  # (B, head, ci, cj)
    bad = (e > 1e-3).nonzero()
    print("bad items:", bad.shape[0], "of", e.numel())
    items = ((bad[:, 0] * lo + bad[:, 2]) * lo + bad[:, 3]) * 4 + bad[:, 1]
    sms = 148
    print("item ids (first 40):", items[:40].tolist())
    print("item seq within CTA (first 40):", (items[:40] // sms).tolist())
    print("CTA ids (first 40):", (items[:40] % sms).tolist())
    # which tiles (rows of the cell) are bad for the first bad item
    b0, hd, ci, cj = bad[0].tolist()
    cell = err[b0, hd * (C // 4):(hd + 1) * (C // 4), ci * r:(ci + 1) * r, cj * r:(cj + 1) * r].amax(dim=0)
    print("first bad item: rows with error:", (cell.amax(dim=1) > 1e-3).nonzero().flatten().tolist())
