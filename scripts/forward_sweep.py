"""The reference's own forward benchmark (test/forward_speed.py + test/forward_memory.py) run on
naf_b200 for the NAF rows it publishes in test/test_results.json (A100-40GB, B=1, fp32, NAF()
defaults: K=9, D=256, 4 heads, random weights).  Same protocol: 5 warm-ups, then 10 forwards timed one
by one with CUDA events, `torch.cuda.empty_cache()` before every forward (allocator cost included),
`torch.no_grad()`; peak memory = `torch.cuda.max_memory_allocated()` after one forward.

    python scripts/forward_sweep.py            # prints one line per published row + a JSON summary
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import naf_b200

# (factor, embed_dim, img_size, lr_size, ratio, published A100 forward ms, published peak MB)
# from the reference's test/test_results.json (NAF entries)
PUBLISHED = [
    ("ratio", 384, 448, 28, 2, 39.513908, 604.98),
    ("ratio", 384, 448, 28, 4, 40.170496, 604.98),
    ("ratio", 384, 448, 28, 8, 42.510439, 604.98),
    ("ratio", 384, 448, 28, 16, 56.241766, 1786.52),
    ("ratio", 384, 448, 28, 32, 267.944962, 7101.49),
    ("embed_dim", 128, 448, 28, 16, 50.801562, 1197.75),
    ("embed_dim", 384, 448, 28, 16, 60.107775, 1786.52),
    ("embed_dim", 768, 448, 28, 16, 88.825344, 2669.67),
    ("embed_dim", 1024, 448, 28, 16, 104.489368, 3258.43),
]
NUM_RUNS = 10


def main():
    dev = torch.device("cuda", 0)
    rows = []
    for factor, C, img_size, lr, ratio, pub_ms, pub_mb in PUBLISHED:
        torch.manual_seed(0)
        model = naf_b200.ModelWrapper(name="NAF", embed_dim=C, ratio=ratio).eval().to(dev)
        img = torch.randn(1, 3, img_size, img_size, device=dev)
        feats = torch.randn(1, C, lr, lr, device=dev)
        size = (ratio * lr, ratio * lr)
        with torch.no_grad():
            for _ in range(5):
                torch.cuda.empty_cache()
                _ = model(img, feats, size)
            total = 0.0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(NUM_RUNS):
                torch.cuda.empty_cache()
                torch.cuda.synchronize()
                e0.record()
                _ = model(img, feats, size)
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
            ms = total / NUM_RUNS
            # steady state (no empty_cache, back to back) for reference
            torch.cuda.synchronize()
            e0.record()
            for _ in range(NUM_RUNS):
                _ = model(img, feats, size)
            e1.record()
            torch.cuda.synchronize()
            ms_steady = e0.elapsed_time(e1) / NUM_RUNS
            # the same forward replayed as a CUDA graph (naf_b200.GraphedNAF)
            fast = naf_b200.GraphedNAF(model.model)
            for _ in range(3):
                _ = fast(img, feats, size)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(NUM_RUNS):
                _ = fast(img, feats, size)
            e1.record()
            torch.cuda.synchronize()
            ms_graph = e0.elapsed_time(e1) / NUM_RUNS
            del _, fast
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            out = model(img, feats, size)
            torch.cuda.synchronize()
            peak_mb = torch.cuda.max_memory_allocated() / 1024 ** 2
            del out
        mpix = size[0] * size[1] / 1e6
        rows.append(dict(factor=factor, embed_dim=C, img_size=img_size, lr_size=lr, ratio=ratio,
                         forward_ms=round(ms, 4), forward_ms_steady=round(ms_steady, 4), forward_ms_graph=round(ms_graph, 4), peak_mb=round(peak_mb, 1),
                         mpix_s=round(mpix / ms * 1e3, 2), published_a100_ms=pub_ms, published_a100_peak_mb=pub_mb,
                         speedup_vs_published_a100=round(pub_ms / ms, 1)))
        print(f"{factor:9s} C={C:4d} img={img_size} lr={lr} x{ratio:<2d}: {ms:8.3f} ms (steady {ms_steady:7.3f}, graph {ms_graph:7.3f})  "
              f"{mpix / ms * 1e3:8.1f} Mpix/s  peak {peak_mb:7.1f} MB | published A100: {pub_ms:8.2f} ms, {pub_mb:7.1f} MB "
              f"-> x{pub_ms / ms:.1f}", flush=True)
    print(json.dumps({"protocol": "reference test/forward_speed.py (B=1, fp32, K=9, 5 warm-up + 10 timed, empty_cache "
                                  "before every forward)", "device": torch.cuda.get_device_name(0), "rows": rows}))


if __name__ == "__main__":
    main()
