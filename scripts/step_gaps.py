"""Is the C2 step launch-bound?  CPU enqueue time per step vs device time per step.
   python scripts/step_gaps.py [workload]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
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
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = model(image, feats, (to, to))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f"{wl}: CPU enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, device {e0.elapsed_time(e1) / n:.3f} ms/step, wall {1e3 * (t2 - t0) / n:.3f} ms/step")
# the same with one step at a time (sync between steps): exposes per-step launch latency
with torch.no_grad():
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        out = model(image, feats, (to, to))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
print("single steps (ms):", " ".join(f"{t:.3f}" for t in ts))
