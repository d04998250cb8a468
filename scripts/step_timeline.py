"""GPU timeline of NAF.forward steps from the profiler trace: busy time vs span, and the largest gaps.
   python scripts/step_timeline.py [workload] [steps]"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import naf_b200
from bench import WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
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
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
busy = sum(e["dur"] for e in ev)
print(f"{wl}: {steps} steps, {len(ev)} GPU ops; span {span / steps / 1e3:.3f} ms/step, busy {busy / steps / 1e3:.3f} ms/step, idle {(span - busy) / steps / 1e3:.3f} ms/step")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    gaps.append((g, a["name"][:50], b["name"][:50]))
for g, a, b in sorted(gaps, reverse=True)[:12]:
    print(f"  gap {g:8.1f} us  after {a:50s} before {b}")
