"""Pins the config-5 part of oracle/model_oracle.py against the REFERENCE (run in the build container,
where /root/reference exists) and writes tests/golden/lstm_golden.pt:

  * ViT_LSTM (reference class, src/models/vln_model.py:163-250) two recurrent steps, B=3, L=7, with the
    vision model replaced by an identity stub (the trunk has its own fixtures) -> outputs / states
  * NavCMTAgent.move_view_corners + get_direction (reference functions, src/xview_lstm/agent.py) on random poses
    incl. poses that leave the map at each stage -> corners / headings

    python tests/golden/make_lstm_golden.py

Shims: `.cuda()` no-op; stub modules for shapely / tensorboardX; transformers.ViTFeatureExtractor.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model_oracle as mo  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
sys.path.insert(0, "/root/reference/src")
for name in ("shapely", "shapely.geometry", "shapely.ops", "tensorboardX"):
    m = types.ModuleType(name)
    for attr in ("Point", "Polygon", "LineString", "MultiPoint", "SummaryWriter", "nearest_points"):
        setattr(m, attr, object)
    sys.modules[name] = m
import transformers  # noqa: E402
transformers.ViTFeatureExtractor = object
np.int = int
np.mat = np.asmatrix
from models.vln_model import ViT_LSTM as RefViTLSTM  # noqa: E402


def close(a, b, tol, what):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: err {err} ref {ref}"


def lstm_case(out):
    torch.manual_seed(3)

    class Ident(torch.nn.Module):
        def forward(self, x):
            return x

    net = RefViTLSTM(types.SimpleNamespace(), Ident())
    net.eval()
    B, L = 3, 7
    feat1 = torch.randn(B, 512, 7, 7) * 0.5
    feat2 = torch.randn(B, 512, 7, 7) * 0.5
    cls_hidden = torch.relu(torch.randn(B, 49))
    lang = torch.randn(B, L, 768)
    d1 = torch.tensor([[10], [200], [355]])
    d2 = torch.tensor([[90], [45], [0]])
    with torch.no_grad():
        h, c, hh, cc, o1, s1 = net(d1, feat1, None, cls_hidden, lang)
        h2, c2, hh2, cc2, o2, s2 = net(d2, feat2, None, cls_hidden, lang, h, c, hh, cc)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        r1 = mo.vit_lstm_step(sd, feat1.view(B, 512, 49), d1, cls_hidden, lang)
        r2 = mo.vit_lstm_step(sd, feat2.view(B, 512, 49), d2, cls_hidden, lang, r1[:4])
    for a, b, n in zip(r1[:5], (h, c, hh, cc, o1), "h c hh cc out".split()):
        close(a, b, 1e-5, "step1 " + n)
    for a, b, n in zip(r2[:5], (h2, c2, hh2, cc2, o2), "h c hh cc out".split()):
        close(a, b, 1e-5, "step2 " + n)
    up = torch.nn.functional.interpolate(r2[5].view(-1, 1, 8, 8), size=(224, 224), mode="bilinear", align_corners=False)
    close(up, s2, 1e-5, "saliency")
    out["lstm"] = dict(sd_small={k: v for k, v in sd.items() if v.numel() <= 4096},
                       sd_recipe="torch.manual_seed(3); ViT_LSTM(args, Identity) default init (make_lstm_golden.py)",
                       keys=sorted(sd.keys()), feat1=feat1, feat2=feat2, cls_hidden=cls_hidden, lang=lang, d1=d1, d2=d2,
                       out1=o1, out2=o2, h2=h2, c2=c2, hh2=hh2, cc2=cc2, h_sali2=r2[5], sal2_sub=s2[:, :, ::16, ::16].clone())
    print("ViT_LSTM ok", o1[0].tolist())


def move_case(out):
    # the reference agent class without its __init__ (BERT download, cv2 windows ...)
    import importlib
    agent_mod = importlib.import_module("xview_lstm.agent")
    RefAgent = agent_mod.NavCMTAgent
    a = RefAgent.__new__(RefAgent)
    rng = np.random.default_rng(5)
    cases = []
    bl, tr = (40.0, -75.0), (40.02, -74.98)
    for i in range(200):
        ctr = np.array([40.01, -74.99]) + rng.uniform(-0.0085, 0.0085, size=2)
        half = rng.uniform(0.0004, 0.0018)
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        sq = np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * half          # FL, FR, BR, BL in (lat, lng)
        corners = ctr + sq @ R.T
        angle = int(rng.integers(0, 361))
        dist = float(rng.uniform(0, 0.003))
        alt = int(rng.integers(40, 401))
        in_dir = int(round(agent_mod.get_direction(np.mean(corners, axis=0), (corners[0] + corners[1]) / 2)) % 360)
        if i % 7 == 0:
            in_dir = (in_dir + 30) % 360            # exercises the heading-correction branch
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            new_c, new_d = a.move_view_corners(corners.copy(), angle, dist, alt, bl, tr, in_dir)
        oc, od = mo.move_view_corners(corners.copy(), angle, dist, alt, bl, tr, in_dir)
        assert np.array_equal(np.asarray(new_c), oc) and new_d == od, (i, new_d, od)
        cases.append(dict(corners=corners, angle=angle, dist=dist, alt=alt, in_dir=in_dir, new_corners=np.asarray(new_c),
                          new_dir=float(new_d)))
    kinds = sum(1 for c in cases if np.array_equal(c["corners"], c["new_corners"]))
    out["move"] = dict(bl=bl, tr=tr, cases=cases)
    print("move_view_corners ok: 200 cases,", kinds, "rejected at the zoom stage")


if __name__ == "__main__":
    out = {}
    lstm_case(out)
    move_case(out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lstm_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path))
