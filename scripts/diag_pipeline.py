"""GPU diagnostic: memory formats through the encoder and per-stage device times (C2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import naf_b200
from naf_b200 import ops

dev = torch.device("cuda", 0)
B, C, gi, to, lo, K = 8, 768, 448, 896, 32, 7
torch.manual_seed(0)
m = naf_b200.NAF(kernel_size=K).eval().to(dev)
img = torch.randn(B, 3, gi, gi, device=dev)
feats = torch.randn(B, C, lo, lo, device=dev)


def t(fn, n=5):
    for _ in range(2):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


with torch.no_grad():
    enc = m.image_encoder
    xcl = img.contiguous(memory_format=torch.channels_last)
    ms, a = t(lambda: enc.encoder(xcl)); print("encoder(1x1) ms", ms, a.stride())
    ms, b = t(lambda: enc.sem_encoder(xcl)); print("sem_encoder(3x3) ms", ms, b.stride())
    ms, c = t(lambda: torch.cat([a, b], 1)); print("cat ms", ms, c.stride())
    ms, x = t(lambda: F.adaptive_avg_pool2d(c, (to, to))); print("pool ms", ms, x.stride(), ops.is_pixel_major(x))
    del a, b, c
    tables = enc.rope.axis_tables(to, to)
    ms, xp = t(lambda: ops.as_pixel_major(x)); print("as_pixel_major ms", ms)
    ms, (k, _) = t(lambda: ops.rope_kpool(xp, tables, 4, pooled_hw=(lo, lo))); print("rope_kpool ms", ms, "GB/s", 4*B*256*to*to/ms/1e6)
    ms, v = t(lambda: ops.pack_nhwc(feats)); print("pack V ms", ms)
    ms, o = t(lambda: ops.xattn(xp, k, feats, 4, K, rope_tables=tables)); print("xattn ms", ms)
    del o
    ms, _ = t(lambda: m.image_encoder.guidance(img, (to, to))); print("guidance total ms", ms)
    ms, _ = t(lambda: m(img, feats, (to, to))); print("forward total ms", ms)
    # stage-by-stage formats inside one encoder
    y = xcl
    for i, layer in enumerate(enc.sem_encoder):
        y = layer(y)
        print("sem layer", i, y.stride(), y.is_contiguous(memory_format=torch.channels_last))
