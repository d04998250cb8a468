"""Per-kernel device time of one NAF.forward step (C2 by default) from torch.profiler (CUPTI):
warm caches, kernels back to back -- complements the cold, serialised ncu launch list.
   python scripts/step_breakdown.py [workload] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import naf_b200
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, C, gi, to, lo, K = WORKLOADS[wl]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = naf_b200.NAF(kernel_size=K).eval().to(dev)
image = torch.randn(B, 3, gi, gi, device=dev)
feats = torch.randn(B, C, lo, lo, device=dev)
with torch.no_grad():
    for _ in range(3):
        out = model(image, feats, (to, to))
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            out = model(image, feats, (to, to))
        torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print(f"# {wl}: {steps} steps, {tot / steps:.3f} ms of kernel time per step")
for k, n, ms in sorted(rows, key=lambda r: -r[2]):
    print(f"{ms / steps:9.3f} ms/step {n // steps:4d}x {100 * ms / tot:5.1f}%  {k[:110]}")
