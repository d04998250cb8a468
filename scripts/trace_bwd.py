"""Phase timeline of the tensor-core backward kernel (profiling build NAF_BWD_EXP=32: clock stamps of thread 0
of four CTAs at every phase boundary).  NAF_B200_LIB=scripts/exp/libnaf_bwdtrace.so python scripts/trace_bwd.py"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import naf_b200
from naf_b200 import _lib, ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
for (B, Cn, Ho, h, K) in [(1, 384, 448, 28, 9), (1, 768, 896, 32, 7)]:
    D, n = 256, 4
    q = torch.randn(B, Ho, Ho, D, device=dev).permute(0, 3, 1, 2); k = torch.randn(B, h, h, D, device=dev).permute(0, 3, 1, 2)
    v = torch.randn(B, h, h, Cn, device=dev).permute(0, 3, 1, 2); dout = torch.randn(B, Ho, Ho, Cn, device=dev).permute(0, 3, 1, 2)
    tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(Ho, Ho)
    fn = lambda: ops.xattn_bwd(q, k, v, dout, n, K, rope_tables=tabs, algo=_lib.ALGO_CELL_TC)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    buf = (C.c_longlong * (4 * 128))()
    rc = _lib.load().naf_debug_bwd_trace(buf)
    nchunks = (Cn // n + 63) // 64
    labels = ["start", "prologue"]
    per_tile = ["D0 staged", "B S done", "C softmax"] + [f"D{c} staged" for c in range(1, nchunks)] + ["E chunks done", "F dS", "G dQdK done", "H dQ stored"]
    print(f"== C={Cn} {Ho}<-{h} K={K}: {e0.elapsed_time(e1):.3f} ms (rc {rc}), {nchunks} chunks")
    for s in range(4):
        t = [buf[s * 128 + i] for i in range(128)]
        nz = [x for x in t if x]
        if len(nz) < 3:
            continue
        n_st = len(nz)
        print(f"-- CTA slot {s}: {n_st} stamps, total {nz[-1] - nz[0]} clk")
        names = labels + [f"t{(i // len(per_tile))} {per_tile[i % len(per_tile)]}" for i in range(n_st - 3)] + ["Z end"]
        for i in range(1, n_st):
            print(f"   {names[i] if i < len(names) else i:22s} +{nz[i] - nz[i - 1]:7d}")
