"""CPU: pin the oracle (oracle/naf_oracle.py) against the reference-generated golden fixtures,
against the live reference modules when /root/reference is mounted, and against closed-form
known answers (the reference itself ships no numerical tests: SURVEY.md section 4)."""
import numpy as np
import pytest
import torch

import golden_util as G
from oracle import naf_oracle as O
from oracle import natten_stub, reference_runner

TOL = 2e-6  # fp32 CPU vs fp32 CPU, different summation order at most


@pytest.mark.parametrize("name", G.names("xattn_"))
def test_attention_matches_golden(name):
    c = G.attention_case(name)
    res = O.cross_attention(c["q"], c["k"], c["v"], c["heads"], c["K"], return_weights=c["scores"] is not None)
    out = res[0] if c["scores"] is not None else res
    assert out.shape == c["out"].shape
    assert (out - c["out"]).abs().max().item() <= TOL
    if c["scores"] is not None:
        assert res[1].shape == c["scores"].shape
        assert (res[1] - c["scores"]).abs().max().item() <= 1e-5
    Ho, Wo = c["q"].shape[-2:]
    h, w = c["k"].shape[-2:]
    assert c["dilation"] == (Ho // h, Wo // w)


@pytest.mark.parametrize("name", G.names("rope_"))
def test_rope_matches_golden(name):
    c = G.rope_case(name)
    D = c["x"].shape[1]
    periods = O.rope_periods(D // c["heads"], 100.0)
    assert torch.equal(periods, c["periods"])
    out = O.rope_rotate(c["x"], c["heads"], periods)
    assert torch.equal(out, c["out"])  # the restatement is bit-identical on CPU


@pytest.mark.parametrize("name", G.names("naf_"))
def test_module_tail_matches_golden(name):
    """RoPE -> key pool -> attention downstream of the stored rotated queries."""
    c = G.module_case(name)
    q = c["queries"]
    k = O.key_pool(q, *c["features"].shape[-2:])
    out = O.cross_attention(q, k, c["features"], 4, 7)
    assert (out - c["out"]).abs().max().item() <= TOL


@pytest.mark.skipif(not reference_runner.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("shape", [(9, 4, 7, 32, 64), (12, 3, 9, 16, 256), (13, None, 9, 8, 64)])
def test_oracle_matches_live_reference(shape):
    h, r, K, C, D = shape
    ns = reference_runner.load()
    torch.manual_seed(1)
    Ho = 32 if r is None else h * r
    model = ns.NAF(dim=D, kernel_size=K).eval()
    img, feats = torch.randn(1, 3, Ho, Ho), torch.randn(1, C, h, h)
    with torch.no_grad():
        ref, ref_scores = model(img, feats, (Ho, Ho), return_weights=True)
        x = model.image_encoder.forward_encoder(img, (Ho, Ho))
    out, scores = O.naf_forward(x, feats, 4, 4, K, return_weights=True)
    assert (out - ref).abs().max().item() <= TOL
    assert (scores - ref_scores).abs().max().item() <= 1e-5


# ---------------------------------------------------------------- known-answer tests
def _rand(*shape, seed=0):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))


def test_zero_query_gives_window_mean():
    h, r, K, C = 8, 3, 5, 6
    q = torch.zeros(1, 16, h * r, h * r)
    k, v = _rand(1, 16, h, h), _rand(1, C, h, h, seed=1)
    out = O.cross_attention(q, k, v, 2, K)
    rt, ct = O.tap_tables(h * r, h * r, h, h, K)
    y, x = 7, 20
    win = v[0][:, rt[y].astype(np.int64)][:, :, ct[x].astype(np.int64)]
    assert torch.allclose(out[0, :, y, x], win.mean(dim=(1, 2)), atol=1e-6)


def test_constant_values_pass_through():
    q, k = _rand(1, 32, 12, 12), _rand(1, 32, 6, 6, seed=1)
    v = torch.full((1, 8, 6, 6), 3.25)
    out = O.cross_attention(q, k, v, 4, 3)
    assert torch.allclose(out, torch.full_like(out, 3.25), atol=1e-6)


def test_window_covering_whole_map_is_dense_attention():
    h, r, n = 5, 2, 2
    q, k, v = _rand(1, 8, h * r, h * r), _rand(1, 8, h, h, seed=1), _rand(1, 6, h, h, seed=2)
    out = O.cross_attention(q, k, v, n, h)
    qh = q.reshape(1, n, 4, -1).permute(0, 1, 3, 2)
    kh = k.reshape(1, n, 4, -1)
    vh = v.reshape(1, n, 3, -1).permute(0, 1, 3, 2)
    dense = torch.softmax(qh @ kh * 4 ** -0.5, dim=-1) @ vh  # (1,n,HW,dv)
    dense = dense.permute(0, 1, 3, 2).reshape(1, 6, h * r, h * r)
    assert torch.allclose(out, dense, atol=2e-6)


def test_ratio_one_kernel_one_is_identity():
    q, k, v = _rand(1, 4, 6, 7), _rand(1, 4, 6, 7, seed=1), _rand(1, 3, 6, 7, seed=2)
    out = O.cross_attention(q, k, v, 1, 1)
    assert torch.allclose(out, v, atol=1e-7)


def test_peaky_logits_select_one_cell():
    h, r, K = 6, 2, 3
    k = torch.zeros(1, 4, h, h)
    k[0, 0, 2, 3] = 1.0
    q = torch.zeros(1, 4, h * r, h * r)
    q[0, 0] = 200.0
    v = _rand(1, 5, h, h)
    out = O.cross_attention(q, k, v, 1, K)
    # pixel (4,6) is in cell (2,3): its window contains (2,3) -> picks v[:,2,3]
    assert torch.allclose(out[0, :, 4, 6], v[0, :, 2, 3], atol=1e-5)


def test_border_cells_share_the_clamped_window():
    h, r, K = 9, 4, 7
    rt, _ = O.tap_tables(h * r, h * r, h, h, K)
    for cell in range(K // 2 + 1):  # cells 0..3 all see rows 0..6
        assert (rt[cell * r:(cell + 1) * r] == np.arange(K)).all()
    assert (rt[-1] == np.arange(h - K, h)).all()


def test_non_integer_ratio_duplicates_taps():
    rt = O.axis_tap_table(32, 13, 9)
    assert rt[0].tolist() == [0, 1, 1, 2, 3, 4, 5, 5, 6]


def test_scores_tap_order_is_row_major():
    h, r, K = 5, 2, 3
    q, k, v = _rand(1, 4, h * r, h * r), _rand(1, 4, h, h, seed=1), _rand(1, 2, h, h, seed=2)
    _, s = O.cross_attention(q, k, v, 1, K, return_weights=True)
    rt, ct = O.tap_tables(h * r, h * r, h, h, K)
    y, x, t, u = 4, 5, 2, 0
    want = (q[0, :, y, x] * k[0, :, rt[y, t], ct[x, u]]).sum() * 4 ** -0.5
    assert abs(s[0, 0, y, x, t * K + u].item() - want.item()) < 1e-6


def test_head_permutation_equivariance():
    n, dq, dv = 2, 4, 3
    q, k, v = _rand(1, n * dq, 8, 8), _rand(1, n * dq, 4, 4, seed=1), _rand(1, n * dv, 4, 4, seed=2)
    out = O.cross_attention(q, k, v, n, 3)
    sw = lambda t, d: torch.cat([t[:, d:], t[:, :d]], dim=1)
    out_sw = O.cross_attention(sw(q, dq), sw(k, dq), sw(v, dv), n, 3)
    assert torch.equal(sw(out, dv), out_sw)


def test_stub_and_oracle_window_rules_agree():
    for L, K, d in [(36, 7, 4), (32, 9, 2), (45, 3, 4), (20, 15, 1), (37 * 28, 11, 28)]:
        a = natten_stub.axis_taps(L, K, d).numpy()
        b = np.array([[O._window_start(i, L, K, d) + t * d for t in range(K)] for i in range(L)])
        assert (a == b).all()


def test_encoder_resolution_rule():
    assert O.image_encoder_resolution(448, 448, 56, 56) == (224, 224)
    assert O.image_encoder_resolution(448, 448, 112, 112) == (448, 448)
    assert O.image_encoder_resolution(224, 224, 448, 448) == (224, 224)


def test_per_pixel_rope_equals_the_map_rotation():
    """`rope_rotate_pixels` (used by the full-size GPU spot checks) == `rope_rotate` on the same pixels."""
    import numpy as np
    x = torch.from_numpy(np.random.RandomState(3).standard_normal((1, 256, 9, 13)).astype(np.float32))
    per = O.rope_periods(64)
    full = O.rope_rotate(x, 4, per)
    ys, xs = [0, 8, 4, 2], [0, 12, 7, 11]
    got = O.rope_rotate_pixels(torch.stack([x[0, :, y, xx] for y, xx in zip(ys, xs)]), ys, xs, 9, 13, 4, per)
    want = torch.stack([full[0, :, y, xx] for y, xx in zip(ys, xs)])
    assert torch.equal(got, want)


@pytest.mark.parametrize("mode", ["separate", "min", "max"])
@pytest.mark.parametrize("train", [False, True])
def test_rope_coordinates_match_live_reference(mode, train):
    """naf_b200.RoPE.axis_coords (pure torch, per axis) == the reference's create_coordinate
    (src/layers/rope.py:84-126) for the three normalisations and, with the same seed, for the train-time
    shift / jitter / rescale augmentations (same torch RNG calls in the same order)."""
    if not reference_runner.available():
        pytest.skip("reference tree not available")
    import naf_b200
    ns = reference_runner.load()
    kw = dict(num_heads=2, base=100.0, normalize_coords=mode, shift_coords=0.1 if train else None,
              jitter_coords=1.5 if train else None, rescale_coords=2.0)
    ref = ns.RoPE(64, **kw)
    ours = naf_b200.RoPE(64, **kw)
    ref.train(train)
    ours.train(train)
    H, W = 7, 11
    torch.manual_seed(123)
    want = ref.create_coordinate(H=H, W=W).view(H, W, 2)
    torch.manual_seed(123)
    cy, cx = ours.axis_coords(H, W)
    assert torch.equal(want[:, 0, 0], cy) and torch.equal(want[0, :, 1], cx)
    assert torch.equal(want[:, :, 0], cy[:, None].expand(H, W)) and torch.equal(want[:, :, 1], cx[None, :].expand(H, W))


@pytest.mark.parametrize("name", G.names("bwd_xattn_"))
def test_oracle_gradients_match_reference_golden(name):
    """The fp64 autograd oracle used as the gradcheck reference of the CUDA backward == gradients of the
    unmodified reference modules (fp32) on the committed fixtures."""
    c = G.bwd_attention_case(name)
    dq, dk, dv = O.cross_attention_grads(c["q"], c["k"], c["v"], c["dout"], c["heads"], c["K"])
    for got, want in ((dq, c["dq"]), (dk, c["dk"]), (dv, c["dv"])):
        assert (got.float() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("name", G.names("rect_xattn_"))
def test_rectangular_windows_match_reference_golden(name):
    """NATTEN's kernel_size=(kh, kw) (the reference passes its tuple through, src/layers/attentions.py:20,24): the
    oracle's forward, scores (tap order t_h*kw + t_w) and fp64 gradients against fixtures from the unmodified reference."""
    c = G.rect_attention_case(name)
    out, scores = O.cross_attention(c["q"], c["k"], c["v"], c["heads"], c["K"], return_weights=True)
    assert scores.shape[-1] == c["K"][0] * c["K"][1]
    assert (out - c["out"]).abs().max().item() <= TOL
    assert (scores - c["scores"]).abs().max().item() <= 1e-5
    dq, dk, dv = O.cross_attention_grads(c["q"], c["k"], c["v"], c["dout"], c["heads"], c["K"])
    for got, want in ((dq, c["dq"]), (dk, c["dk"]), (dv, c["dv"])):
        assert (got.float() - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
