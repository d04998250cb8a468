"""TEST INFRASTRUCTURE ONLY -- golden GRADIENT fixtures (tests/golden/bwd_*.npz) produced by autograd
through the UNMODIFIED reference modules (NATTEN replaced by oracle/natten_stub.py, whose two functionals
are plain differentiable torch code).  Run in the build container:

    python -m oracle.gen_golden_bwd

Operator level (reference CrossAttention.forward, src/layers/attentions.py:53-75): dq, dk, dv for a seeded
dout.  Module level (reference NAF.forward, src/model/naf.py:104-116, the call train.py:127,136 back-props
through): d features, d image and the gradients of every parameter (full tensors for the small ones, the
leading slice + sum / abs-sum for the 3x3 conv weights).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import reference_runner
from oracle.gen_golden import OUT, seeded_normal

# name, B, D, heads, C, (Ho,Wo), (h,w), K, q-gain
ATTENTION_CASES = [
    ("bwd_xattn_int_r4_k7", 1, 256, 4, 32, (36, 36), (9, 9), 7, 1.0),
    ("bwd_xattn_int_r3_k5_peaky", 2, 128, 2, 24, (24, 30), (8, 10), 5, 3.0),
    ("bwd_xattn_nonint_32_13_k9", 2, 64, 4, 16, (32, 32), (13, 13), 9, 1.0),     # training-style ratio
    ("bwd_xattn_denoise_r1_k15_c3", 1, 96, 1, 3, (20, 22), (20, 22), 15, 1.0),   # denoising shape
]

# name, (Hi,Wi), (Ho,Wo), (h,w), C -- NAF(dim=128, kernel_size=7), state of naf_dim128_k7_state.npz
MODULE_CASES = [
    ("bwd_naf_same_res", (36, 36), (36, 36), (9, 9), 32),
    ("bwd_naf_train_like", (56, 56), (32, 32), (13, 13), 16),   # image pooled DOWN to the target, 32 <- 13
    ("bwd_naf_replicate_up", (18, 18), (36, 36), (9, 9), 16),
]


def main() -> None:
    ns = reference_runner.load()
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, B, D, n, C, (Ho, Wo), (h, w), K, gain) in enumerate(ATTENTION_CASES):
        seed = 5000 + 10 * idx
        q = (seeded_normal(seed, B, D, Ho, Wo) * gain).requires_grad_(True)
        k = seeded_normal(seed + 1, B, D, h, w).requires_grad_(True)
        v = seeded_normal(seed + 2, B, C, h, w).requires_grad_(True)
        dout = seeded_normal(seed + 3, B, C, Ho, Wo)
        mod = ns.CrossAttention(dim=D, num_heads=n, kernel_size=(K, K))
        out = mod(q, k, v, None)
        out.backward(dout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed), gain=np.float32(gain),
                            q_shape=np.asarray(q.shape), k_shape=np.asarray(k.shape), v_shape=np.asarray(v.shape),
                            heads=np.int64(n), kernel_size=np.int64(K), out=out.detach().contiguous().numpy(),
                            dq=q.grad.numpy(), dk=k.grad.numpy(), dv=v.grad.numpy())
        print("wrote", name, "|dq|max", float(q.grad.abs().max()), "|dk|max", float(k.grad.abs().max()),
              "|dv|max", float(v.grad.abs().max()))

    state = np.load(os.path.join(OUT, "naf_dim128_k7_state.npz"))
    model = ns.NAF(dim=128, kernel_size=7)
    model.load_state_dict({k_: torch.from_numpy(state[k_]) for k_ in state.files})
    model.train()        # what train.py runs; RoPE's train-time augmentations are off (None) in NAF's defaults? see below
    # NAF passes rope_rescale=2.0 (src/model/naf.py:83): in training mode the reference RoPE rescales the
    # coordinates RANDOMLY (src/layers/rope.py:108-124).  The fixtures pin the deterministic part of the
    # backward, so they are taken in eval mode (GroupNorm has no running statistics: same arithmetic).
    model.eval()
    for idx, (name, (Hi, Wi), (Ho, Wo), (h, w), C) in enumerate(MODULE_CASES):
        seed = 6000 + 10 * idx
        img = seeded_normal(seed, 1, 3, Hi, Wi).requires_grad_(True)
        feats = seeded_normal(seed + 1, 1, C, h, w).requires_grad_(True)
        dout = seeded_normal(seed + 2, 1, C, Ho, Wo)
        model.zero_grad()
        out = model(img, feats, (Ho, Wo))
        out.backward(dout)
        d = dict(seed=np.int64(seed), image_shape=np.asarray(img.shape), features_shape=np.asarray(feats.shape),
                 output_size=np.asarray([Ho, Wo], dtype=np.int64), out=out.detach().contiguous().numpy(),
                 dimage=img.grad.numpy(), dfeatures=feats.grad.numpy())
        for pname, p in model.named_parameters():
            g = p.grad
            key = "grad__" + pname.replace(".", "__")
            if g.numel() <= 4096:
                d[key] = g.numpy()
            else:
                d[key + "__head"] = g.reshape(-1)[:4096].numpy().copy()
                d[key + "__sums"] = np.asarray([g.double().sum().item(), g.double().abs().sum().item()])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print("wrote", name, tuple(out.shape), "|dfeat|max", float(feats.grad.abs().max()),
              "|dimg|max", float(img.grad.abs().max()))


if __name__ == "__main__":
    main()
