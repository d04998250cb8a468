"""TEST INFRASTRUCTURE ONLY -- generate the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference):

    cd /root/repo && python -m oracle.gen_golden

Every fixture is produced by the UNMODIFIED reference modules (loaded by
`oracle/reference_runner.py`, NATTEN replaced by `oracle/natten_stub.py`) on seeded CPU fp32
inputs; inputs AND outputs are stored so that nothing depends on RNG reproducibility.  The GPU
parity tests compare the CUDA path against these files on the GPU box, where /root/reference
does not exist.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import reference_runner


def seeded_normal(seed: int, *shape) -> torch.Tensor:
    """Inputs come from numpy's frozen legacy MT19937 stream so that the fixtures only need to
    store the seed (tests/golden_util.py regenerates the identical tensors)."""
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, B, D, heads, C, (Ho,Wo), (h,w), K, q-gain ("peaky" softmax when > 1), return_weights
ATTENTION_CASES = [
    ("xattn_int_r4_k7", 1, 256, 4, 32, (36, 36), (9, 9), 7, 1.0, True),
    ("xattn_int_r3_k11", 2, 256, 4, 64, (36, 36), (12, 12), 11, 3.0, False),
    ("xattn_int_r5_k7_c384", 1, 256, 4, 384, (40, 40), (8, 8), 7, 2.0, False),
    ("xattn_nonsquare_r3x4_k5", 1, 128, 2, 24, (24, 40), (8, 10), 5, 1.0, True),
    ("xattn_nonint_32_13_k9", 2, 64, 4, 16, (32, 32), (13, 13), 9, 1.0, True),
    ("xattn_nonint_30_7_k3", 1, 64, 2, 10, (30, 45), (7, 11), 3, 4.0, False),
    ("xattn_denoise_r1_k15_c3", 1, 96, 1, 3, (20, 22), (20, 22), 15, 1.0, True),
    ("xattn_k_equals_h", 1, 64, 4, 8, (14, 14), (7, 7), 7, 1.0, False),
]

# name, dim, heads, (H, W)
ROPE_CASES = [
    ("rope_d256_n4", 256, 4, (12, 20)),
    ("rope_d64_n1", 64, 1, (7, 5)),
    ("rope_d96_n2", 96, 2, (9, 9)),
]

# name, (Hi,Wi), (Ho,Wo), (h,w), C -- all share one NAF(dim=128, kernel_size=7) instance
MODULE_CASES = [
    ("naf_same_res", (36, 36), (36, 36), (9, 9), 32),
    ("naf_pool_down", (48, 60), (24, 30), (8, 10), 16),
    ("naf_replicate_up", (18, 18), (36, 36), (9, 9), 16),
    ("naf_cap_4x", (100, 100), (20, 20), (10, 10), 8),
]


def main() -> None:
    ns = reference_runner.load()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(20261017)

    for idx, (name, B, D, n, C, (Ho, Wo), (h, w), K, gain, rw) in enumerate(ATTENTION_CASES):
        seed = 1000 + 10 * idx
        q = seeded_normal(seed, B, D, Ho, Wo) * gain
        k = seeded_normal(seed + 1, B, D, h, w)
        v = seeded_normal(seed + 2, B, C, h, w)
        mod = ns.CrossAttention(dim=D, num_heads=n, kernel_size=(K, K)).eval()
        with torch.no_grad():
            if rw:
                out, scores = mod(q, k, v, None, return_weights=True)
            else:
                out, scores = mod(q, k, v, None), None
        d = dict(seed=np.int64(seed), gain=np.float32(gain), q_shape=np.asarray(q.shape),
                 k_shape=np.asarray(k.shape), v_shape=np.asarray(v.shape),
                 out=out.contiguous().numpy(), heads=np.int64(n), kernel_size=np.int64(K),
                 dilation=np.asarray(mod.dilation, dtype=np.int64))
        if scores is not None:
            d["scores"] = scores.contiguous().numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print("wrote", name, tuple(out.shape))

    for idx, (name, D, n, (H, W)) in enumerate(ROPE_CASES):
        seed = 2000 + idx
        x = seeded_normal(seed, 2, D, H, W)
        mod = ns.RoPE(embed_dim=D, num_heads=n, base=100.0, rescale_coords=2.0).eval()
        with torch.no_grad():
            y = mod(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed), x_shape=np.asarray(x.shape),
                            out=y.contiguous().numpy(), periods=mod.periods.numpy(), heads=np.int64(n))
        print("wrote", name)

    torch.manual_seed(7)
    model = ns.NAF(dim=128, kernel_size=7).eval()
    sd = {k_: v_.numpy() for k_, v_ in model.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "naf_dim128_k7_state.npz"), **sd)
    for idx, (name, (Hi, Wi), (Ho, Wo), (h, w), C) in enumerate(MODULE_CASES):
        seed = 3000 + 10 * idx
        img = seeded_normal(seed, 1, 3, Hi, Wi)
        feats = seeded_normal(seed + 1, 1, C, h, w)
        with torch.no_grad():
            out = model(img, feats, (Ho, Wo))
            x = model.image_encoder(img, output_size=(Ho, Wo))  # rotated guidance = queries
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed),
                            image_shape=np.asarray(img.shape), features_shape=np.asarray(feats.shape),
                            out=out.contiguous().numpy(), queries=x.contiguous().numpy(),
                            output_size=np.asarray([Ho, Wo], dtype=np.int64))
        print("wrote", name, tuple(out.shape))

    # structural pins the reference itself publishes (test/test_results.json:63,255)
    full = ns.NAF().eval()
    n_params = sum(p.numel() for p in full.parameters())
    keys = sorted(full.state_dict().keys())
    shapes = {k_: list(v_.shape) for k_, v_ in full.state_dict().items()}
    import json

    with open(os.path.join(OUT, "naf_default_structure.json"), "w") as fh:
        json.dump({"n_params": n_params, "keys": keys, "shapes": shapes,
                   "kernel_size": list(full.upsampler.kernel_size)}, fh, indent=1)
    print("default NAF params:", n_params, "keys:", len(keys))


if __name__ == "__main__":
    main()
