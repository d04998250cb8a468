"""CPU: host-side mirror of the reference interface -- constructor keywords, state_dict keys and
shapes (37 keys, 662,528 parameters: reference test/test_results.json:63,255), the encoder
resolution rule, RoPE tables, error behaviour."""
import json
import os

import pytest
import torch

import golden_util as G
import naf_b200
from naf_b200 import ops
from naf_b200.model.naf import ImageEncoder
from oracle import naf_oracle as O


def test_default_state_dict_matches_reference_structure():
    ref = json.load(open(os.path.join(G.GOLDEN, "naf_default_structure.json")))
    m = naf_b200.NAF()
    sd = m.state_dict()
    assert sorted(sd) == ref["keys"] and len(sd) == 37
    assert all(list(v.shape) == ref["shapes"][k] for k, v in sd.items())
    assert sum(p.numel() for p in m.parameters()) == ref["n_params"] == 662528
    assert list(m.upsampler.kernel_size) == ref["kernel_size"] == [9, 9]


def test_reference_checkpoint_loads_strict():
    m = naf_b200.NAF(dim=128, kernel_size=7)
    missing = m.load_state_dict(G.module_state(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # extra reference kwargs are swallowed like the reference's **kwargs (config/model/naf.yaml)
    naf_b200.NAF(dim=64, use_semencoder=True, name="naf")


def test_encoder_on_cpu_matches_golden_queries_up_to_rope():
    """The conv encoder is library code: with the reference weights it must reproduce the
    reference's pooled guidance map; rotating that with the oracle gives the stored queries."""
    m = naf_b200.NAF(dim=128, kernel_size=7).eval()
    m.load_state_dict(G.module_state())
    for name in G.names("naf_"):
        c = G.module_case(name)
        with torch.no_grad():
            x = m.image_encoder.guidance(c["image"], c["output_size"])
        q = O.rope_rotate(x, 4, O.rope_periods(32, 100.0))
        assert (q - c["queries"]).abs().max().item() < 1e-5, name


def test_encoder_resolution_rule():
    f = ImageEncoder.encoder_resolution
    assert f((448, 448), (56, 56)) == (224, 224)
    assert f((448, 448), (112, 112)) == (448, 448)
    assert f((224, 224), (448, 448)) == (224, 224)
    assert f((100, 100), (20, 20)) == (80, 80)
    assert f((400, 100), (20, 40)) == (80, 80)  # the reference's mixed H/W min (naf.py:42-45)


def test_rope_periods_and_tables_bit_exact():
    rope = naf_b200.RoPE(256, num_heads=4, base=100.0, rescale_coords=2.0).eval()
    assert torch.equal(rope.periods, O.rope_periods(64, 100.0))
    H, W = 12, 20
    cy, sy, cx, sx = ops.rope_axis_tables(H, W, rope.periods)
    cos, sin = O.rope_tables(H, W, rope.periods)  # (HW, 64) as the reference builds them
    cos, sin = cos.view(H, W, 64), sin.view(H, W, 64)
    assert torch.equal(cos[:, 0, :16], cy) and torch.equal(sin[:, 0, :16], sy)
    assert torch.equal(cos[0, :, 16:32], cx) and torch.equal(sin[0, :, 16:32], sx)
    assert torch.equal(cos[..., :32], cos[..., 32:])  # tile(2)


def test_constructor_errors_match_reference():
    with pytest.raises(AssertionError):
        naf_b200.CrossAttention(dim=30, num_heads=4)
    with pytest.raises(AssertionError):
        naf_b200.RoPE(30, num_heads=4)
    with pytest.raises(ValueError):
        naf_b200.RoPE(64, num_heads=4, base=None)
    with pytest.raises(ValueError):
        naf_b200.RoPE(64, num_heads=4, base=10.0, min_period=1.0, max_period=2.0)
    naf_b200.RoPE(64, num_heads=4, base=None, min_period=1.0, max_period=2.0)
    with pytest.raises(ValueError):
        naf_b200.ModelWrapper("JAFAR")
    assert naf_b200.CrossAttention(256, 4, (9, 9)).scale == 0.125


def test_training_mode_augmentation_and_coordinate_cache():
    """Train mode draws the rescale augmentation (reference src/layers/rope.py:119-124); like the
    reference (:159-161) the coordinates are cached per map size, so the same size re-uses them."""
    import torch
    rope = naf_b200.RoPE(64, num_heads=4, rescale_coords=2.0).train()
    torch.manual_seed(5)
    a = rope.axis_tables(4, 6)
    assert rope.axis_tables(4, 6) is a                      # cached per (H, W)
    plain = naf_b200.RoPE(64, num_heads=4, rescale_coords=2.0).eval().axis_tables(4, 6)
    assert not torch.equal(a[0], plain[0])                  # augmented coordinates differ from eval ones
    with pytest.raises(ValueError):
        naf_b200.RoPE(64, num_heads=4, normalize_coords="diagonal").axis_tables(4, 4)
    for mode in ("min", "max"):
        t = naf_b200.RoPE(64, num_heads=4, normalize_coords=mode).eval().axis_tables(4, 6)
        assert t[0].shape == (4, 4) and t[2].shape == (6, 4)


def test_default_width_encoder_on_cpu_plus_oracle_matches_reference_golden():
    """Module tree (torch CPU) + oracle tail == the unmodified reference at the default width
    (dim=256), the fixtures the tensor-core encoder is held to on the GPU."""
    from oracle import naf_oracle as O

    m = naf_b200.NAF(kernel_size=7).eval()
    m.load_state_dict(G.default_state(), strict=True)
    for name in G.names("naf256_"):
        c = G.default_case(name)
        with torch.no_grad():
            x = m.image_encoder.guidance(c["image"], c["output_size"])
        out = O.naf_forward(x.contiguous(), c["features"], 4, 4, 7)
        assert (out - c["out"]).abs().max().item() <= 2e-5, name
