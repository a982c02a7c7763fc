"""CPU: oracle/model_oracle.py against the fixtures the REFERENCE modules produced
(tests/golden/make_model_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "model_golden.pt"), weights_only=False)


def _close(a, b, tol):
    return (a - b).abs().max().item() <= tol * max(b.abs().max().item(), 1e-6)


def test_darknet_forward_backward_matches_reference(golden):
    g = golden["darknet"]
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in g["sd"].items()}
    y = mo.darknet_forward(g["x"], sd, g["cfg"], train=True, update_running=True)
    assert _close(y, g["y"], 1e-5)
    y.backward(g["dy"])
    for n, gr in g["grads"].items():
        assert _close(sd[n].grad, gr, 1e-4), n
    for k, v in g["running_after"].items():
        assert _close(sd[k].detach(), v, 1e-5), k


def test_masks_bit_exact(golden):
    e = golden["et"]
    assert torch.equal(mo.attention_mask(12, 3), e["mask_attn"])
    assert torch.equal(mo.mask_pad(e["lenths"], 12), e["mask_pad"])
    m = mo.attention_mask(3, 2)
    ninf = float("-inf")
    assert m.tolist() == [[0, 0, 0, ninf, ninf, ninf, ninf], [0, 0, 0, ninf, ninf, ninf, ninf],
                          [0, 0, 0, ninf, ninf, ninf, ninf], [0, 0, 0, 0, ninf, 0, ninf], [0, 0, 0, 0, 0, 0, 0],
                          [0, 0, 0, 0, ninf, 0, ninf], [0, 0, 0, 0, 0, 0, 0]]


def test_yolov3_trunk_cfg_shape():
    defs = mo.parse_cfg_text(mo.yolov3_trunk_cfg())[1:]
    assert len(defs) == 80
    assert sum(d["type"] == "convolutional" for d in defs) == 57
    assert sum(d["type"] == "shortcut" for d in defs) == 23
    assert defs[-1]["filters"] == "512"


def test_postprocess_waypoints_integers():
    out = np.array([[2.0, 1.0, 1.5, 0.7], [0.0, -0.5, 0.25, 0.2], [-0.3, 0.3, -1.0, 0.5]], dtype=np.float32)
    ang, dist, alt, stop = mo.postprocess_waypoints(out, np.array([2e-3, 1e-3, 4e-3]))
    assert alt.tolist() == [400, 130, 40]
    assert stop.tolist() == [True, False, False]
    assert ang.tolist() == [63, 180, 315]


def test_nss_matches_reference_agent(golden_dir):
    """``mo.nss`` against the values the reference's own ``NavCMTAgent.NSS`` (src/xview_et/agent.py:256-270)
    returned for the same maps (tests/golden/make_nss_golden.py), for all three nss_r variants."""
    g = torch.load(os.path.join(golden_dir, "nss_golden.pt"), weights_only=False)
    B = g["h_sali"].shape[0]
    sal = torch.nn.functional.interpolate(g["h_sali"].view(B, 1, 8, 8), size=(224, 224), mode="bilinear",
                                          align_corners=False)
    fix = torch.from_numpy(np.unpackbits(g["fix_packed"])[:B * 224 * 224].reshape(B, 224, 224).astype(np.float64))
    for nss_r in (0, 1, -1):
        per = torch.stack([mo.nss(sal[i], fix[i], nss_r) for i in range(B)])
        assert torch.allclose(per, g["per_sample"][nss_r], rtol=1e-9, atol=0), nss_r
        assert torch.allclose(mo.nss(sal.view(B, 224, 224), fix, nss_r), g["batch"][nss_r], rtol=1e-9, atol=0)
