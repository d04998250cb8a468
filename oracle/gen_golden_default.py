"""TEST INFRASTRUCTURE ONLY -- golden fixtures of the DEFAULT-width reference model.

    python -m oracle.gen_golden_default

`NAF(kernel_size=7)` with the reference defaults (dim=256: two 128-channel conv branches, the shape
the tensor-core encoder kernels implement) run UNMODIFIED on CPU fp32 (NATTEN replaced by
`oracle/natten_stub.py`), on seeded inputs.  Stored: the state_dict, the output, and every 16th
channel of the rotated guidance map (the full map would be several MB per case).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import reference_runner
from oracle.gen_golden import OUT, seeded_normal

# name, B, (Hi,Wi), (Ho,Wo), (h,w), C
CASES = [
    ("naf256_same_res", 2, (48, 56), (48, 56), (8, 8), 16),       # whole tiles, batch 2
    ("naf256_replicate_up", 1, (24, 32), (48, 64), (8, 8), 16),   # rep 2, 1.5 tiles high
    ("naf256_ragged_nonint", 1, (37, 29), (37, 29), (9, 9), 8),   # ragged tiles, non-integer ratio
]
QSTEP = 16


def main() -> None:
    ns = reference_runner.load()
    torch.manual_seed(11)
    model = ns.NAF(kernel_size=7).eval()
    # non-trivial GroupNorm affine parameters and biases (the default init is weight=1, bias=0)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "norm" in name:
                p.add_(0.25 * torch.randn(p.shape, generator=g))
    sd = {k_: v_.numpy() for k_, v_ in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "naf256_k7_state.npz"), **sd)
    for idx, (name, B, (Hi, Wi), (Ho, Wo), (h, w), C) in enumerate(CASES):
        seed = 4000 + 10 * idx
        img = seeded_normal(seed, B, 3, Hi, Wi)
        feats = seeded_normal(seed + 1, B, C, h, w)
        with torch.no_grad():
            out = model(img, feats, (Ho, Wo))
            q = model.image_encoder(img, output_size=(Ho, Wo))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed),
                            image_shape=np.asarray(img.shape), features_shape=np.asarray(feats.shape),
                            out=out.contiguous().numpy(), queries_sub=q[:, ::QSTEP].contiguous().numpy(),
                            qstep=np.int64(QSTEP), output_size=np.asarray([Ho, Wo], dtype=np.int64))
        print("wrote", name, tuple(out.shape))


if __name__ == "__main__":
    main()
