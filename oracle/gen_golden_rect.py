"""TEST INFRASTRUCTURE ONLY -- golden fixtures for RECTANGULAR windows (tests/golden/rect_xattn_*.npz): the
UNMODIFIED reference CrossAttention (src/layers/attentions.py:32-75) with kernel_size=(kh, kw), which it hands to
NATTEN as is (:20,24), NATTEN replaced by oracle/natten_stub.py; forward output, pre-softmax scores
(return_weights=True) and the gradients dq, dk, dv by autograd through the reference.  Run in the build container:

    python -m oracle.gen_golden_rect
"""
from __future__ import annotations

import os

import numpy as np

from oracle import reference_runner
from oracle.gen_golden import OUT, seeded_normal

# name, B, D, heads, C, (Ho,Wo), (h,w), (kh,kw), q-gain
CASES = [
    ("rect_xattn_int_r4_k7x5", 1, 128, 2, 24, (32, 40), (8, 10), (7, 5), 1.0),
    ("rect_xattn_int_r3x5_k3x9_peaky", 2, 64, 4, 16, (27, 45), (9, 9), (3, 9), 3.0),
    ("rect_xattn_nonint_30x45_k5x3", 1, 64, 2, 10, (30, 45), (7, 11), (5, 3), 1.0),
]


def main() -> None:
    ns = reference_runner.load()
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, B, D, n, C, (Ho, Wo), (h, w), K, gain) in enumerate(CASES):
        seed = 7000 + 10 * idx
        q = (seeded_normal(seed, B, D, Ho, Wo) * gain).requires_grad_(True)
        k = seeded_normal(seed + 1, B, D, h, w).requires_grad_(True)
        v = seeded_normal(seed + 2, B, C, h, w).requires_grad_(True)
        dout = seeded_normal(seed + 3, B, C, Ho, Wo)
        mod = ns.CrossAttention(dim=D, num_heads=n, kernel_size=K)
        out, scores = mod(q, k, v, None, return_weights=True)
        out.backward(dout)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed), gain=np.float32(gain),
                            q_shape=np.asarray(q.shape), k_shape=np.asarray(k.shape), v_shape=np.asarray(v.shape),
                            heads=np.int64(n), kernel_size=np.asarray(K, dtype=np.int64),
                            dilation=np.asarray(mod.dilation, dtype=np.int64),
                            out=out.detach().contiguous().numpy(), scores=scores.detach().contiguous().numpy(),
                            dq=q.grad.numpy(), dk=k.grad.numpy(), dv=v.grad.numpy())
        print("wrote", name, tuple(out.shape), tuple(scores.shape))


if __name__ == "__main__":
    main()
