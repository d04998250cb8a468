"""Union-window tensor-core kernel (NAF_ALGO_UNION_TC) against the generic fp32 kernel and the CPU oracle, with
timings: the reference's training shape (32 <- 13..19, K = 9: utils/training.py:28-50), the denoising shape
(ratio 1, C = 3, one head, K up to 15: denoising.py:209-213,436-451), small integer ratios.

    python scripts/check_union.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from naf_b200 import _lib, ops
from oracle import naf_oracle as O

# name, B, D, heads, C, (Ho, Wo), (h, w), K, gain, rope
CASES = [
    ("nonint 32<-13 K9 dq16 dv4", 2, 64, 4, 16, (32, 32), (13, 13), 9, 1.0, False),
    ("nonint 30x45<-7x11 K3", 1, 64, 2, 10, (30, 45), (7, 11), 3, 4.0, False),
    ("denoise r1 K15 dq96 C3", 1, 96, 1, 3, (20, 22), (20, 22), 15, 1.0, False),
    ("int r4 K7", 1, 256, 4, 32, (36, 36), (9, 9), 7, 1.0, False),
    ("int r2 K9 rope", 1, 256, 4, 384, (56, 56), (28, 28), 9, 2.0, True),
    ("train 32<-13 K9 C384", 8, 256, 4, 384, (32, 32), (13, 13), 9, 2.0, True),
    ("train 32<-19 K9 C768", 8, 256, 4, 768, (32, 32), (19, 19), 9, 2.0, True),
    ("train 37<-16 K9 C1024", 4, 256, 4, 1024, (37, 37), (16, 16), 9, 1.0, True),
    ("denoise r1 K15 dq256 C3 256^2", 2, 256, 1, 3, (256, 256), (256, 256), 15, 3.0, True),
    ("denoise r1 K7 dq96 C3 128^2", 4, 96, 1, 3, (128, 128), (128, 128), 7, 3.0, True),
    ("denoise r1 K15 dq128 heads1-4", 1, 128, 1, 3, (96, 80), (96, 80), 15, 2.0, True),
    # shapes AUTO gives to the fp32 SIMT cell kernel
    ("int r14 K7 dq32 dv16", 2, 128, 4, 64, (448, 448), (32, 32), 7, 2.0, True),
    ("int r14 K7 dq64 dv10", 2, 256, 4, 40, (224, 224), (16, 16), 7, 2.0, True),
    ("int r14 K13 dq64 dv96", 2, 256, 4, 384, (224, 224), (16, 16), 13, 2.0, True),
    ("int r4 K9 dq64 dv96", 2, 256, 4, 384, (112, 112), (28, 28), 9, 2.0, True),
    ("int r8 K9 dq64 dv96", 1, 256, 4, 384, (224, 224), (28, 28), 9, 2.0, True),
]


def rnd(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda", 0)
    worst = 0.0
    for idx, (name, B, D, n, C, (Ho, Wo), (h, w), K, gain, rope) in enumerate(CASES):
        q = (rnd(10 * idx, B, D, Ho, Wo) * gain).to(dev)
        k = rnd(10 * idx + 1, B, D, h, w).to(dev)
        v = rnd(10 * idx + 2, B, C, h, w).to(dev)
        tabs = None
        if rope:
            import naf_b200
            r = naf_b200.RoPE(D, num_heads=n, base=100.0, rescale_coords=2.0).eval().to(dev)
            tabs = r.axis_tables(Ho, Wo)
        run = lambda algo: ops.xattn(q, k, v, n, K, rope_tables=tabs, algo=algo)
        ref = run(_lib.ALGO_GENERIC)
        got = run(_lib.ALGO_UNION_TC)
        err = (got - ref).abs().max().item()
        worst = max(worst, err)
        line = f"{name:34s} max|union - generic| = {err:.2e}"
        if B * Ho * Wo * C <= 1 << 22 and not rope:
            want = O.cross_attention(q.cpu(), k.cpu(), v.cpu(), n, K)
            line += f"   max|union - oracle| = {(got.cpu() - want).abs().max().item():.2e}"
        t_g = timed(lambda: run(_lib.ALGO_GENERIC))
        t_u = timed(lambda: run(_lib.ALGO_UNION_TC))
        line += f"   generic {t_g * 1e3:8.1f} us   union {t_u * 1e3:8.1f} us"
        for algo in (_lib.ALGO_CELL_SIMT, _lib.ALGO_CELL_TCWS, _lib.ALGO_CELL_TMA):
            try:
                t_a = timed(lambda: run(algo))
                line += f"   {_lib.ALGO_NAMES[algo]} {t_a * 1e3:8.1f} us"
            except Exception:
                pass
        line += f"   auto -> {ops.select_algo(q.shape, v.shape, n, K, rope)}"
        print(line, flush=True)
    print("worst", worst)
    assert worst <= 1e-3


if __name__ == "__main__":
    main()
