"""Host <-> device copy bandwidth from pinned memory, with and without the GPU-local CPU affinity (NVML)."""
import os, subprocess, sys, time
import torch

def bw(n_bytes=4 << 30, reps=3, tag=""):
    d = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    out = []
    for direction in ("d2h", "h2d"):
        for _ in range(2):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        out.append(f"{direction} {n_bytes * reps / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
    print(tag, ", ".join(out), flush=True)
    del d, h

print("affinity now:", sorted(os.sched_getaffinity(0)))
for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c"],
            ["sh", "-c", "lscpu | grep -i numa"], ["sh", "-c", "cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null"]):
    try:
        print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout)
    except Exception as e:
        print(cmd, "failed:", e)
torch.cuda.init()
bw(tag="default affinity:")
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    try:
        print("ideal cpu mask words:", list(pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)))
    except Exception as e:
        print("GetCpuAffinity failed:", e)
    pynvml.nvmlDeviceSetCpuAffinity(hnd)
    print("affinity after nvmlDeviceSetCpuAffinity:", sorted(os.sched_getaffinity(0)))
    bw(tag="GPU-local affinity:")
except Exception as e:
    print("nvml affinity failed:", e)
# try every single NUMA-ish block of CPUs
cpus = sorted(os.sched_getaffinity(0))
