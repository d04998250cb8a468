"""CPU: the product's host tap tables (naf_b200/taps.py) are bit-identical to the oracle's and to
ATen's nearest-exact index map; the closed form used on the device equals the table for every
integer ratio (all five BASELINE configs included)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from naf_b200 import taps
from oracle import naf_oracle as O

CONFIG_AXES = [(224, 16, 7), (896, 32, 7), (1036, 37, 11), (1344, 24, 7), (2048, 32, 7), (448, 28, 9)]
ODD_AXES = [(32, 13, 9), (30, 7, 3), (45, 11, 3), (100, 33, 5), (20, 20, 15), (50, 7, 7), (64, 9, 7), (23, 5, 3)]


@pytest.mark.parametrize("out_len,in_len", [(a, b) for a, b, _ in CONFIG_AXES + ODD_AXES] + [(7, 3), (1000, 999), (4096, 37)])
def test_nearest_exact_matches_aten(out_len, in_len):
    src = torch.arange(in_len, dtype=torch.float32).view(1, 1, in_len, 1)
    got = F.interpolate(src, size=(out_len, 1), mode="nearest-exact").view(-1).to(torch.int64).numpy()
    assert (taps.nearest_exact_source(out_len, in_len) == got).all()
    assert (O.nearest_exact_index(out_len, in_len) == got).all()


@pytest.mark.parametrize("out_len,in_len,K", CONFIG_AXES + ODD_AXES)
def test_axis_taps_bit_exact_vs_oracle(out_len, in_len, K):
    a = taps.axis_taps(out_len, in_len, K)
    b = O.axis_tap_table(out_len, in_len, K)
    assert a.dtype == np.int32 and a.shape == (out_len, K)
    assert (a == b).all()
    assert a.min() >= 0 and a.max() < in_len


@pytest.mark.parametrize("out_len,in_len,K", CONFIG_AXES + [(36, 9, 7), (36, 12, 11), (14, 7, 7), (20, 20, 15), (24, 8, 5)])
def test_closed_form_equals_table_for_integer_ratios(out_len, in_len, K):
    assert taps.integer_ratio(out_len, in_len)
    assert (taps.closed_form_axis_taps(out_len, in_len, K) == taps.axis_taps(out_len, in_len, K)).all()


def test_tap_tables_none_for_integer_ratios():
    assert taps.tap_tables(896, 896, 32, 32, 7) == (None, None)
    rt, ct = taps.tap_tables(32, 48, 13, 16, 9)
    assert rt.shape == (32, 9) and ct.shape == (48, 9)


@pytest.mark.parametrize("args", [(36, 9, 4), (36, 9, 0), (20, 7, 11), (5, 9, 3)])
def test_invalid_windows_raise(args):
    with pytest.raises(ValueError):
        taps.axis_taps(*args)


# ---- the union-window formulation of naf_xattn_union.cu, on the host ---------------------------------
@pytest.mark.parametrize("Ho,Wo,h,w,K", [(32, 32, 13, 13, 9), (30, 45, 7, 11, 3), (20, 22, 20, 22, 15), (36, 36, 9, 9, 7), (37, 41, 16, 18, 9)])
def test_union_window_multiplicity_softmax_equals_tap_softmax(Ho, Wo, h, w, K):
    """Dense attention of an 8 x 16 pixel tile against the rectangle of cells its windows touch, weighted by the
    per-pixel tap multiplicities mr[y][i] * mc[x][j], equals softmax over the K*K taps (duplicated taps included):
    the identity the union kernel computes, checked with the oracle's tables and its attention."""
    rt, ct = O.tap_tables(Ho, Wo, h, w, K)
    rs = np.random.RandomState(3)
    D, C = 16, 5
    q = torch.from_numpy(rs.standard_normal((1, D, Ho, Wo)).astype(np.float32)) * 2
    k = torch.from_numpy(rs.standard_normal((1, D, h, w)).astype(np.float32))
    v = torch.from_numpy(rs.standard_normal((1, C, h, w)).astype(np.float32))
    want = O.cross_attention(q, k, v, 1, K)[0]                       # (C, Ho, Wo)
    scale = D ** -0.5
    worst_span = 0
    for ty0 in range(0, Ho, 8):
        for tx0 in range(0, Wo, 16):
            ys, xs = np.arange(ty0, min(ty0 + 8, Ho)), np.arange(tx0, min(tx0 + 16, Wo))
            i0, i1, j0, j1 = rt[ys].min(), rt[ys].max(), ct[xs].min(), ct[xs].max()
            RH, RW = i1 - i0 + 1, j1 - j0 + 1
            worst_span = max(worst_span, RH, RW)
            ku = k[0, :, i0:i1 + 1, j0:j1 + 1].reshape(D, -1)       # (D, |U|)
            vu = v[0, :, i0:i1 + 1, j0:j1 + 1].reshape(C, -1)
            for y in ys:
                mr = np.bincount(rt[y] - i0, minlength=RH)
                for x in xs:
                    mc = np.bincount(ct[x] - j0, minlength=RW)
                    m = torch.from_numpy(np.outer(mr, mc).reshape(-1).astype(np.float32))
                    s = (q[0, :, y, x] @ ku) * scale
                    e = m * torch.exp(s - s[m > 0].max())
                    got = (vu @ e) / e.sum()
                    assert (got - want[:, y, x]).abs().max().item() <= 2e-6, (y, x)
    assert worst_span <= 32      # the kernel's RMAX: tables built from NATTEN windows stay inside it


@pytest.mark.parametrize("shape", [(32, 40, 8, 10, (7, 5)), (27, 45, 9, 9, (3, 9)), (30, 45, 7, 11, (5, 3)), (896, 896, 32, 32, (7, 11))])
def test_rectangular_window_tables_bit_exact_vs_oracle(shape):
    """kernel_size=(kh, kw): row table (Ho, kh), column table (Wo, kw), each the square rule of its own axis."""
    Ho, Wo, h, w, K = shape
    rt_o, ct_o = O.tap_tables(Ho, Wo, h, w, K)
    assert rt_o.shape == (Ho, K[0]) and ct_o.shape == (Wo, K[1])
    assert (taps.axis_taps(Ho, h, K[0]) == rt_o).all() and (taps.axis_taps(Wo, w, K[1]) == ct_o).all()
    rt, ct = taps.tap_tables(Ho, Wo, h, w, K)
    if Ho % h == 0 and Wo % w == 0:
        assert rt is None and ct is None
        assert (taps.closed_form_axis_taps(Ho, h, K[0]) == rt_o).all() and (taps.closed_form_axis_taps(Wo, w, K[1]) == ct_o).all()
    else:
        assert (rt == rt_o).all() and (ct == ct_o).all()
