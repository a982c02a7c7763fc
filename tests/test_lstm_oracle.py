"""CPU: the config-5 oracle (ViT_LSTM step, move_view_corners, waypoint_step) against the golden vectors the
REFERENCE produced (tests/golden/make_lstm_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "lstm_golden.pt"), weights_only=False)


def _state_dict(g):
    """Rebuild the fixture's weights from its seed recipe with stock torch modules in the reference's
    construction order (vln_model.py:170-209), pinned by the stored small tensors."""
    from torch import nn
    torch.manual_seed(3)
    mods = {}
    mods["direction_embedding"] = nn.Linear(2, 32)
    mods["pos_embedding"] = nn.Linear(2, 32)
    for n in ("attention_layer_lang", "attention_layer_vision_lang"):
        mods[n + ".linear_in"] = nn.Linear(768, 768, bias=False)
        mods[n + ".linear_out"] = nn.Linear(1536, 768, bias=False)
    mods["attention_layer_vision.linear_in"] = nn.Linear(49, 49, bias=False)
    mods["attention_layer_vision.linear_out"] = nn.Linear(98, 49, bias=False)
    mods["vision_lstm"] = nn.LSTMCell(49, 576)
    mods["direct_lstm"] = nn.LSTMCell(32, 192)
    mods["decoder_2_action_full.0"], mods["decoder_2_action_full.3"], mods["decoder_2_action_full.6"] = \
        nn.Linear(768, 256), nn.Linear(256, 32), nn.Linear(32, 4)
    mods["fc.0"], mods["fc.3"] = nn.Linear(49, 128), nn.Linear(128, 64)
    sd = {}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    assert sorted(sd.keys()) == g["keys"]
    for k, v in g["sd_small"].items():
        assert torch.equal(sd[k], v), k
    return sd


def test_vit_lstm_two_steps_vs_reference_golden(golden):
    g = golden["lstm"]
    sd = _state_dict(g)
    B = g["feat1"].shape[0]
    with torch.no_grad():
        r1 = mo.vit_lstm_step(sd, g["feat1"].view(B, 512, 49), g["d1"], g["cls_hidden"], g["lang"])
        r2 = mo.vit_lstm_step(sd, g["feat2"].view(B, 512, 49), g["d2"], g["cls_hidden"], g["lang"], r1[:4])
    assert torch.allclose(r1[4], g["out1"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(r2[4], g["out2"], rtol=1e-5, atol=1e-6)
    for a, k in zip(r2[:4], ("h2", "c2", "hh2", "cc2")):
        assert torch.allclose(a, g[k], rtol=1e-5, atol=1e-6), k


def test_move_view_corners_vs_reference_golden(golden):
    g = golden["move"]
    for c in g["cases"]:
        nc, nd = mo.move_view_corners(c["corners"].copy(), c["angle"], c["dist"], c["alt"], g["bl"], g["tr"], c["in_dir"])
        assert np.array_equal(nc, c["new_corners"]) and float(nd) == c["new_dir"]


def test_waypoint_step_semantics():
    """Stop / last-step samples keep their pose and become ended; the others move."""
    rng = np.random.default_rng(0)
    B = 6
    ctr = np.array([40.01, -74.99])
    sq = np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * 0.001
    corners = np.stack([ctr + sq for _ in range(B)])
    bounds = np.tile(np.array([40.0, -75.0, 40.02, -74.98]), (B, 1))
    out = rng.normal(size=(B, 4)).astype(np.float32)
    out[:, 3] = [0.9, 0.1, 0.2, 0.3, 0.0, 0.26]
    nc, nd, ended, ang, alt, dist = mo.waypoint_step(out, corners, bounds, np.zeros(B), np.zeros(B, bool), 0.25, False)
    assert ended.tolist() == [True, False, False, True, False, True]
    for i in range(B):
        assert np.array_equal(nc[i], corners[i]) == bool(ended[i])
    _, _, ended2, *_ = mo.waypoint_step(out, corners, bounds, np.zeros(B), np.zeros(B, bool), 0.25, True)
    assert ended2.all()
