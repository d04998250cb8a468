"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of NAF's cross-scale neighbourhood
attention forward.  It is the checker; it is never the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``naf_b200/`` does.

What it restates (reference file:line it follows):

* ``tap_tables``        -- the composition of NATTEN's shifted dilated window
                           (third-party NATTEN 0.17.x `get_window_start`, called at
                           /root/reference/src/layers/attentions.py:20,24 with
                           ``dilation = (Ho//h, Wo//w)``, :56) with ``F.interpolate(mode=
                           "nearest-exact")`` of K and V (:48-51,60-61).  Integer tables
                           ``row_tap[Ho][K]``, ``col_tap[Wo][K]`` -- the objects the CUDA path
                           must reproduce bit-exactly.
* ``rope_tables`` / ``rope_rotate`` -- /root/reference/src/layers/rope.py:84-105 (coords, eval
                           branch, "separate" normalisation), :128-135 (periods), :137-153
                           (angles, tile(2), cos/sin), :15-34 (rotate-half pairing).
* ``key_pool``          -- /root/reference/src/model/naf.py:63-69 (adaptive_avg_pool2d of the
                           ROTATED map to the feature resolution).
* ``cross_attention``   -- /root/reference/src/layers/attentions.py:16-29,46-75: scale
                           ``(dim//heads)**-0.5`` applied after QK (:46,21), softmax over the
                           K*K taps in row-major tap order (:23), weighted sum of V (:24),
                           optional return of the *scaled pre-softmax* scores (:27-28).
* ``naf_forward``       -- /root/reference/src/model/naf.py:104-116 given the pooled,
                           un-rotated guidance map ``x`` (the encoder is ATen/cuDNN library
                           code and is simply *used*, not restated).

PARITY STATUS: the reference holds no golden vectors for this path (its test/ directory only
times things).  This oracle is pinned against outputs of the UNMODIFIED reference modules run
in the build container with `oracle/natten_stub.py` standing in for NATTEN
(`oracle/gen_golden.py` -> `tests/golden/*.npz`).  The NATTEN window rule itself cannot be
checked against a real NATTEN binary offline: at that single third-party boundary parity is
"unpinned" (see DESIGN.md).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- integer tables
def _window_start(i: int, L: int, K: int, d: int) -> int:
    ns = K // 2
    if d <= 1:
        s = max(i - ns, 0)
        if i + ns >= L:
            s += L - i - ns - 1
        return s
    if i - ns * d < 0:
        return i % d
    if i + ns * d >= L:
        m = i % d
        a = (L // d) * d
        b = L - a
        return (L - b + m - 2 * ns * d) if m < b else (a + m - K * d)
    return i - ns * d


def nearest_exact_index(out_len: int, in_len: int) -> np.ndarray:
    """Source index of every destination index for F.interpolate(mode="nearest-exact").

    ATen: src = min(floor((dst + 0.5) * scale), in_len - 1) with scale = float32(in)/out."""
    scale = np.float32(in_len) / np.float32(out_len)
    dst = np.arange(out_len, dtype=np.float32)
    src = np.floor((dst + np.float32(0.5)) * scale).astype(np.int64)
    return np.minimum(src, in_len - 1)


def axis_tap_table(out_len: int, in_len: int, K: int) -> np.ndarray:
    """(out_len, K) int32: low-res index of tap t of target index i along one axis."""
    if K < 1 or K % 2 == 0:
        raise ValueError("kernel_size must be odd and positive")
    d = out_len // in_len
    if d < 1 or K * d > out_len:
        raise ValueError(f"kernel_size*dilation ({K}*{d}) does not fit axis of length {out_len}")
    near = nearest_exact_index(out_len, in_len)
    tab = np.empty((out_len, K), dtype=np.int32)
    for i in range(out_len):
        s = _window_start(i, out_len, K, d)
        for t in range(K):
            tab[i, t] = near[s + t * d]
    return tab


def tap_tables(Ho: int, Wo: int, h: int, w: int, K) -> tuple[np.ndarray, np.ndarray]:
    kh, kw = (K, K) if isinstance(K, int) else (int(K[0]), int(K[1]))
    return axis_tap_table(Ho, h, kh), axis_tap_table(Wo, w, kw)


# ----------------------------------------------------------------------------------- RoPE
def rope_periods(d_head: int, base: float = 100.0, dtype=torch.float32) -> torch.Tensor:
    return base ** (2 * torch.arange(d_head // 4, dtype=dtype) / (d_head // 2))


def rope_tables(H: int, W: int, periods: torch.Tensor, dtype=torch.float32):
    """cos, sin of shape (H*W, D_head), exactly as RoPE.rotate builds them (eval mode)."""
    ch = torch.arange(0.5, H, dtype=dtype) / H
    cw = torch.arange(0.5, W, dtype=dtype) / W
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
    coords = 2.0 * coords - 1.0
    ang = 2 * math.pi * coords[:, :, None] / periods.to(dtype)[None, None, :]
    ang = ang.flatten(1, 2).tile(2)
    return torch.cos(ang), torch.sin(ang)


def rope_rotate(x: torch.Tensor, heads: int, periods: torch.Tensor) -> torch.Tensor:
    """x (B, D, H, W) -> rotated (B, D, H, W); channel c belongs to rope-head c // (D/heads)."""
    B, D, H, W = x.shape
    dh = D // heads
    assert D % (4 * heads) == 0
    cos, sin = rope_tables(H, W, periods, dtype=x.dtype)  # (HW, dh)
    xv = x.reshape(B, heads, dh, H * W).permute(0, 1, 3, 2)  # (B,n,HW,dh)
    a, b = xv[..., : dh // 2], xv[..., dh // 2 :]
    rot = torch.cat([-b, a], dim=-1)
    out = xv * cos + rot * sin
    return out.permute(0, 1, 3, 2).reshape(B, D, H, W)


def rope_rotate_pixels(xv: torch.Tensor, ys, xs, H: int, W: int, heads: int, periods: torch.Tensor) -> torch.Tensor:
    """Rotation of individual pixels: xv (N, D) holds the un-rotated vectors of the pixels
    (ys[i], xs[i]) of an H x W map -> (N, D).  Same formulas as `rope_tables` / `rope_rotate`
    (/root/reference/src/layers/rope.py:99-105,139-153) evaluated only where asked, so that the
    full-size spot checks need no (H*W, D_head) table."""
    N, D = xv.shape
    dh = D // heads
    dt = xv.dtype
    ys = torch.as_tensor(ys, dtype=dt)
    xs = torch.as_tensor(xs, dtype=dt)
    coords = torch.stack([(ys + 0.5) / H, (xs + 0.5) / W], dim=-1)      # arange(0.5, H) / H
    coords = 2.0 * coords - 1.0
    ang = 2 * math.pi * coords[:, :, None] / periods.to(dt)[None, None, :]
    ang = ang.flatten(1, 2).tile(2)                                      # (N, dh)
    cos, sin = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
    v = xv.reshape(N, heads, dh)
    a, b = v[..., : dh // 2], v[..., dh // 2 :]
    rot = torch.cat([-b, a], dim=-1)
    return (v * cos + rot * sin).reshape(N, D)


def key_pool(q: torch.Tensor, h: int, w: int) -> torch.Tensor:
    return F.adaptive_avg_pool2d(q, output_size=(h, w))


# ------------------------------------------------------------------------------ attention
def cross_attention(q, k, v, num_heads: int, kernel_size, return_weights: bool = False,
                    rows_per_chunk: int = 8):
    """q (B,D,Ho,Wo) rotated guidance, k (B,D,h,w), v (B,C,h,w) -> (B,C,Ho,Wo).

    Direct low-resolution-window form: tap (t,u) of pixel (y,x) reads low-res cell
    (row_tap[y][t], col_tap[x][u]).  No high-resolution copy of K or V is ever made."""
    B, D, Ho, Wo = q.shape
    _, C, h, w = v.shape
    assert k.shape == (B, D, h, w)
    assert D % num_heads == 0, "dim must be divisible by num_heads"
    assert C % num_heads == 0
    n = num_heads
    dq, dv = D // n, C // n
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    rt_np, ct_np = tap_tables(Ho, Wo, h, w, (kh, kw))
    rt = torch.from_numpy(rt_np.astype(np.int64))
    ct = torch.from_numpy(ct_np.astype(np.int64))
    scale = dq ** -0.5
    k = k.to(q.dtype)
    v = v.to(q.dtype)
    qh = q.reshape(B, n, dq, Ho, Wo).permute(0, 1, 3, 4, 2)  # (B,n,Ho,Wo,dq)
    kk = k.reshape(B, n, dq, h, w).permute(0, 1, 3, 4, 2)    # (B,n,h,w,dq)
    vv = v.reshape(B, n, dv, h, w).permute(0, 1, 3, 4, 2)    # (B,n,h,w,dv)
    out = q.new_empty(B, n, Ho, Wo, dv)
    scores = q.new_empty(B, n, Ho, Wo, kh * kw) if return_weights else None
    for y0 in range(0, Ho, rows_per_chunk):
        y1 = min(Ho, y0 + rows_per_chunk)
        r = rt[y0:y1]
        kn = kk[:, :, r[:, :, None, None], ct[None, None, :, :], :]  # (B,n,hc,kh,Wo,kw,dq)
        s = (qh[:, :, y0:y1, None, :, None, :] * kn).sum(-1) * scale  # (B,n,hc,kh,Wo,kw)
        s = s.permute(0, 1, 2, 4, 3, 5).reshape(B, n, y1 - y0, Wo, kh * kw)
        if return_weights:
            scores[:, :, y0:y1] = s
        p = s.softmax(dim=-1).reshape(B, n, y1 - y0, Wo, kh, kw).permute(0, 1, 2, 4, 3, 5)
        vn = vv[:, :, r[:, :, None, None], ct[None, None, :, :], :]  # (B,n,hc,kh,Wo,kw,dv)
        out[:, :, y0:y1] = (p[..., None] * vn).sum(dim=(3, 5))
    res = out.permute(0, 1, 4, 2, 3).reshape(B, C, Ho, Wo)
    return (res, scores) if return_weights else res


def naf_forward(x_pooled, features, heads_attn: int, heads_rope: int, kernel_size,
                rope_base: float = 100.0, return_weights: bool = False):
    """NAF.forward downstream of the conv encoder + adaptive pool: RoPE -> K pool -> attention."""
    B, D, Ho, Wo = x_pooled.shape
    periods = rope_periods(D // heads_rope, rope_base, dtype=x_pooled.dtype)
    q = rope_rotate(x_pooled, heads_rope, periods)
    k = key_pool(q, features.shape[-2], features.shape[-1])
    return cross_attention(q, k, features, heads_attn, kernel_size, return_weights)


def image_encoder_resolution(Hi: int, Wi: int, Ho: int, Wo: int) -> tuple[int, int]:
    """Resolution the conv encoder runs at (/root/reference/src/model/naf.py:39-48)."""
    if Hi > 4 * Ho or Wi > 4 * Wo:
        return min(Hi, 4 * Ho, 4 * Wo), min(Wi, 4 * Wo, 4 * Ho)
    return Hi, Wi


# ------------------------------------------------------------------------------ backward
def cross_attention_grads(q, k, v, dout, num_heads: int, kernel_size, dtype=torch.float64):
    """Gradients (dq, dk, dv) of sum(cross_attention(q, k, v) * dout), by torch autograd through this
    file's own forward restatement (evaluated in `dtype`, fp64 by default: the gradcheck reference of
    the CUDA backward).  Follows what autograd does to the reference's legacy_attention
    (/root/reference/src/layers/attentions.py:16-29): softmax backward on the scaled logits, dV as the
    scatter-add of P^T dOut over the shared windows, dK likewise from dS^T Q."""
    qd, kd, vd = (t.detach().to(dtype).requires_grad_(True) for t in (q, k, v))
    out = cross_attention(qd, kd, vd, num_heads, kernel_size)
    out.backward(dout.to(dtype))
    return qd.grad, kd.grad, vd.grad


def naf_forward_grads(x_pooled, features, dout, heads_attn: int, heads_rope: int, kernel_size,
                      rope_base: float = 100.0, dtype=torch.float64):
    """Gradients (dx, dfeatures) of sum(naf_forward(x, features) * dout): the hot path downstream of the
    conv encoder -- dQ through the transposed rotation, dK through the key pooling
    (/root/reference/src/model/naf.py:63-69,108: K is pooled from the ROTATED map, so every pixel also
    receives 1/|bin| of its bins' dK), dV scatter-add."""
    xd, fd = (t.detach().to(dtype).requires_grad_(True) for t in (x_pooled, features))
    out = naf_forward(xd, fd, heads_attn, heads_rope, kernel_size, rope_base)
    out.backward(dout.to(dtype))
    return xd.grad, fd.grad
