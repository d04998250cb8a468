"""Tensor-core backward cell kernel (NAF_ALGO_CELL_TC of naf_xattn_bwd_f32) against the fp64 oracle, the fp32
cell kernel and the generic kernel, with timings on the reference's backward-benchmark shapes
(test/backward_speed.py: B = 1, 448 x 448 target, K = 9).

    python scripts/check_bwd_tc.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import naf_b200
from naf_b200 import _lib, ops
from oracle import naf_oracle as O

# name, B, C, (Ho, Wo), (h, w), K, gain, rope, oracle
CASES = [
    ("r14 K3 dv32", 1, 128, (56, 56), (4, 4), 3, 1.0, False, True),
    ("r8 K7 dv96 peaky", 2, 384, (64, 96), (8, 12), 7, 3.0, False, True),
    ("r12 K9 dv24->generic? dv16", 1, 64, (108, 108), (9, 9), 9, 1.0, False, True),
    ("r14 K11 dv64", 1, 256, (154, 154), (11, 11), 11, 2.0, False, True),
    ("r28 K7 dv192 rope", 1, 768, (224, 224), (8, 8), 7, 2.0, True, False),
    ("sweep C384 r16", 1, 384, (448, 448), (28, 28), 9, 1.0, True, False),
    ("sweep C128 r16", 1, 128, (448, 448), (28, 28), 9, 1.0, True, False),
    ("sweep C768 r16", 1, 768, (448, 448), (28, 28), 9, 1.0, True, False),
    ("sweep C1024 r16", 1, 1024, (448, 448), (28, 28), 9, 1.0, True, False),
    ("sweep C384 r8", 1, 384, (224, 224), (28, 28), 9, 1.0, True, False),
    ("C2-like B2 r28 K7", 2, 768, (896, 896), (32, 32), 7, 1.0, True, False),
]


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b.cpu()).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def main():
    dev = torch.device("cuda", 0)
    worst = 0.0
    D, n = 256, 4
    for idx, (name, B, C, (Ho, Wo), (h, w), K, gain, rope, oracle) in enumerate(CASES):
        q = (rnd(10 * idx, B, D, Ho, Wo) * gain).to(dev)
        k = rnd(10 * idx + 1, B, D, h, w).to(dev)
        v = rnd(10 * idx + 2, B, C, h, w).to(dev)
        dout = rnd(10 * idx + 3, B, C, Ho, Wo).to(dev)
        tabs = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev).axis_tables(Ho, Wo) if rope else None
        run = lambda algo: ops.xattn_bwd(q, k, v, dout, n, K, rope_tables=tabs, algo=algo)
        try:
            got = run(_lib.ALGO_CELL_TC)
        except Exception as e:  # unsupported shape
            print(f"{name:28s} cell_tc: {e}")
            continue
        ref = run(_lib.ALGO_CELL_SIMT)
        errs = [rel(a, b) for a, b in zip(got, ref)]
        line = f"{name:28s} tc vs simt (dq, dk, dv) = " + " ".join(f"{e:.1e}" for e in errs)
        if oracle:
            want = O.cross_attention_grads(q.cpu(), k.cpu(), v.cpu(), dout.cpu(), n, K)
            e64 = [rel(a, b) for a, b in zip(got, want)]
            s64 = [rel(a, b) for a, b in zip(ref, want)]
            line += "   tc vs fp64 = " + " ".join(f"{e:.1e}" for e in e64) + "   simt vs fp64 = " + " ".join(f"{e:.1e}" for e in s64)
            worst = max(worst, *e64)
        worst = max(worst, *errs)
        t_tc = timed(lambda: run(_lib.ALGO_CELL_TC))
        t_s = timed(lambda: run(_lib.ALGO_CELL_SIMT), n=3)
        fwd = timed(lambda: ops.xattn(q, k, v, n, K, rope_tables=tabs))
        gb = 4.0 * B * Ho * Wo * (C + 2 * D) / 1e9
        line += f"   tc {t_tc:7.3f} ms ({gb / t_tc * 1e3:6.0f} GB/s)   simt {t_s:8.3f} ms   fwd {fwd:6.3f} ms"
        print(line, flush=True)
    print("worst", worst)
    assert worst <= 2e-5


if __name__ == "__main__":
    main()
