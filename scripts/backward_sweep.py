"""The reference's own backward benchmark (test/backward_speed.py + test/backward_memory.py) run on naf_b200
for the NAF rows it publishes in test/test_results.json (A100-40GB, B=1, fp32, NAF() defaults: K=9, D=256,
4 heads, random weights).  Same protocol: model + a 1x1 Conv2d(embed_dim, 1) head, SGD over both, 5 warm-up
steps, then 10 steps (forward, loss = head(output).sum(), backward, optimizer step) timed one by one with CUDA
events, `torch.cuda.empty_cache()` around every step; peak memory = max_memory_allocated over one step.

    python scripts/backward_sweep.py          # one line per published row + a JSON summary
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn, optim

import naf_b200

# (factor, embed_dim, img_size, lr_size, ratio, published A100 step ms, published peak MB)
PUBLISHED = [
    ("ratio", 384, 448, 28, 2, 88.292659, 3670.42),
    ("ratio", 384, 448, 28, 4, 102.527796, 3684.27),
    ("ratio", 384, 448, 28, 8, 112.663655, 4028.77),
    ("ratio", 384, 448, 28, 16, 163.075171, 6016.48),
    ("embed_dim", 128, 448, 28, 16, 132.395419, 5139.84),
    ("embed_dim", 768, 448, 28, 16, 201.987173, 7487.63),
    ("embed_dim", 1024, 448, 28, 16, 222.735052, 8468.40),
]
NUM_RUNS = 10


def main():
    dev = torch.device("cuda", 0)
    rows = []
    for factor, C, img_size, lr, ratio, pub_ms, pub_mb in PUBLISHED:
        torch.manual_seed(0)
        model = naf_b200.ModelWrapper(name="NAF", embed_dim=C, ratio=ratio).to(dev)   # train mode, like the reference
        img = torch.randn(1, 3, img_size, img_size, device=dev)
        feats = torch.randn(1, C, lr, lr, device=dev)
        size = (ratio * lr, ratio * lr)
        head = nn.Conv2d(C, 1, 1).to(dev)
        opt = optim.SGD(list(model.parameters()) + list(head.parameters()), lr=0.01)

        def step():
            out = model(img, feats, size)
            loss = head(out).sum()
            loss.backward()
            opt.step()

        for _ in range(5):
            torch.cuda.empty_cache()
            step()
            torch.cuda.empty_cache()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(NUM_RUNS):
            torch.cuda.empty_cache()
            torch.cuda.synchronize()
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            total += e0.elapsed_time(e1)
        ms = total / NUM_RUNS
        torch.cuda.reset_peak_memory_stats()
        step()
        torch.cuda.synchronize()
        mb = torch.cuda.max_memory_allocated() / 2 ** 20
        # the attention path alone: forward + backward of the fused op on a detached guidance map
        x = torch.randn(1, 256, size[0], size[1], device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        f = feats.clone().requires_grad_(True)
        g = torch.randn(1, C, size[0], size[1], device=dev)
        for _ in range(2):
            model.model.upsample_from_guidance(x, f).backward(g)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            model.model.upsample_from_guidance(x, f).backward(g)
        e1.record()
        torch.cuda.synchronize()
        attn_ms = e0.elapsed_time(e1) / 5
        rows.append(dict(factor=factor, embed_dim=C, img_size=img_size, lr_size=lr, ratio=ratio, step_ms=round(ms, 3),
                         published_a100_ms=pub_ms, peak_mb=round(mb, 1), published_peak_mb=pub_mb,
                         attention_fwd_bwd_ms=round(attn_ms, 3)))
        print(f"C={C:5d} ratio={ratio:3d} target={size[0]:4d}: step {ms:8.3f} ms (A100 published {pub_ms:8.2f}), peak {mb:8.1f} MB "
              f"(published {pub_mb:8.1f}); attention path fwd+bwd alone {attn_ms:7.3f} ms", flush=True)
        del model, head, opt
    print(json.dumps({"protocol": "test/backward_speed.py", "rows": rows}))


if __name__ == "__main__":
    main()
