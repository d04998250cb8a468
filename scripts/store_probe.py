"""Store-path micro-benchmark (tests/cuda/store_probe.cu): GB/s of the epilogue's store schemes alone."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from naf_b200.csrc.build import nvcc_path
src = os.path.join(ROOT, "tests", "cuda", "store_probe.cu")
lib = os.path.join(ROOT, "tests", "cuda", "libstore_probe.so")
if not os.path.isfile(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
    subprocess.check_call([nvcc_path(), "-std=c++17", "-O2", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                           "-shared", "-I", os.path.join(ROOT, "naf_b200", "csrc"), "-o", lib, src])
L = C.CDLL(lib)
L.store_probe.restype = C.c_int
L.store_probe.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]
L.store_probe_rw.restype = C.c_int
L.store_probe_rw.argtypes = [C.c_void_p] + [C.c_int] * 11 + [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]
dev = torch.device("cuda", 0)
Ho = Wo = 28 * 7 * 16      # 3136: 112 x 16 cells
Cn, DV, th, tw = 768, 192, 4, 28
out = torch.empty(Ho, Wo, Cn, device=dev)
iters = 200
def run(mode, per_group, slots, grid=148, tag="", DV=DV, th=th):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        rc = L.store_probe(out.data_ptr(), Ho, Wo, Cn, mode, iters, th, tw, DV, per_group, slots, grid, st)
        assert rc == 0, rc
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        L.store_probe(out.data_ptr(), Ho, Wo, Cn, mode, iters, th, tw, DV, per_group, slots, grid, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gb = grid * iters * 7 * th * tw * DV * 4 / 1e9
    print(f"mode {mode} per_group {per_group} slots {slots} grid {grid} {tag}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb:.2f} GB)")
run(0, 3, 2, tag="tensor stores 3 boxes/group, 2 slots (the kernel today)")
run(0, 6, 2, tag="tensor stores 6 boxes/group, 2 slots")
run(0, 6, 1, tag="tensor stores 6 boxes/group, 1 slot")
run(0, 1, 3, tag="tensor stores 1 box/group, 3 slots")
run(0, 2, 3, tag="tensor stores 2 boxes/group, 3 slots")
run(0, 3, 3, tag="tensor stores 3 boxes/group, 3 slots")
run(1, 0, 1, tag="per-thread 1-D bulk copies 384 B")
run(2, 1, 2, tag="one unswizzled 192-channel box per tile, 2 slots")
run(2, 1, 1, tag="one unswizzled 192-channel box per tile, 1 slot")
run(0, 3, 2, grid=74, tag="half the SMs")

# reference: plain contiguous fill of the same tensor (torch elementwise kernel)
for _ in range(2):
    out.fill_(1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out.fill_(2.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"torch fill_ of {out.numel() * 4 / 1e9:.2f} GB: {ms:.3f} ms  {out.numel() * 4 / 1e9 / ms * 1e3:.0f} GB/s")
# whole pixels (all 768 channels contiguous, 3 KB per pixel): does the chunk size matter?
run(2, 1, 1, tag="unswizzled box of 256 channels x 28 x 4 pixels (1 KB contiguous per pixel)", DV=256, th=4)
run(1, 0, 1, tag="per-thread 1-D bulk copies 1536 B (all 768 channels, 2 threads per pixel)", DV=768, th=2)


# ---- the launch's read : write mix without any compute: tensor stores as in the kernel + a paced stream of
# 32-byte query loads (8 KB per tile ~ C2 through rep=2: 7 KB; 28 KB per tile = C2 with x at the target resolution)
rd = torch.randn(512 * 1024 * 1024, device=dev)   # 2 GB read buffer
sink = torch.zeros(1, device=dev)
def run_rw(rbytes, mode=0, per_group=3, slots=2, tag=""):
    st = torch.cuda.current_stream().cuda_stream
    args = (out.data_ptr(), Ho, Wo, Cn, mode, iters, th, tw, DV, per_group, slots, 148, st, rd.data_ptr(), rbytes, rd.numel(), sink.data_ptr())
    for _ in range(2):
        rc = L.store_probe_rw(*args)
        assert rc == 0, rc
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 3
    for _ in range(n):
        L.store_probe_rw(*args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    wgb = 148 * iters * 7 * th * tw * DV * 4 / 1e9
    rgb = 148 * iters * 7 * rbytes / 1e9
    print(f"stores + reads, {rbytes} B read per tile, mode {mode} {tag}: {ms:.3f} ms  written {wgb:.2f} GB + read {rgb:.2f} GB = {(wgb + rgb) / ms * 1e3:.0f} GB/s total")
run_rw(0, tag="(no reads, 384-thread launch)")
run_rw(8192, tag="(C2 mix, x through rep=2)")
run_rw(28672, tag="(C2 mix, x at target resolution)")
run_rw(8192, mode=2, per_group=1, tag="(unswizzled 192-channel boxes)")
run_rw(28672, mode=2, per_group=1, tag="(unswizzled 192-channel boxes)")
