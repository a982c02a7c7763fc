"""Pins ``oracle.model_oracle.nss`` against the REFERENCE's own ``NavCMTAgent.NSS`` (src/xview_et/agent.py:256-270),
run in the build container where /root/reference exists, and writes tests/golden/nss_golden.pt: the 8x8 head
outputs the saliency maps are upsampled from, bit-packed fixation maps, and the reference's NSS value for
nss_r in {0, 1, -1} -- per sample (the training loop calls NSS one sample at a time, agent.py:673-681) and for
the whole batch.

    python tests/golden/make_nss_golden.py

The agent class is used without its __init__ (BERT download, tensorboard ...), as make_lstm_golden.py does for
``move_view_corners``.  Shims: `.cuda()` no-op; stub modules for shapely / tensorboardX;
transformers.ViTFeatureExtractor; np.int / np.mat.
"""
import importlib
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


def main():
    importlib.import_module("xview_lstm.agent")            # first: SURVEY.md §8c shim (4)
    RefAgent = importlib.import_module("xview_et.agent").NavCMTAgent
    agent = RefAgent.__new__(RefAgent)
    torch.manual_seed(7)
    B = 5
    h_sali = torch.randn(B, 64) * 0.7
    sal = torch.nn.functional.interpolate(h_sali.view(B, 1, 8, 8), size=(224, 224), mode="bilinear",
                                          align_corners=False)
    rng = np.random.default_rng(3)
    fix = np.zeros((B, 224, 224), dtype=np.float64)
    for i in range(B):                                       # filled discs, as env.py:224-230 rasterises them
        for _ in range(int(rng.integers(1, 4))):
            cy, cx, r = rng.integers(0, 224), rng.integers(0, 224), rng.integers(5, 60)
            yy, xx = np.ogrid[:224, :224]
            fix[i][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1.0
    fix_t = torch.from_numpy(fix)
    out = dict(h_sali=h_sali, fix_packed=np.packbits(fix > 0), per_sample={}, batch={})
    for nss_r in (0, 1, -1):
        agent.args = types.SimpleNamespace(nss_r=nss_r)
        per = torch.stack([agent.NSS(sal[i], fix_t[i]) for i in range(B)])        # agent.py:679 call shape
        whole = agent.NSS(sal.view(B, 224, 224), fix_t)
        mine = torch.stack([mo.nss(sal[i], fix_t[i], nss_r) for i in range(B)])
        assert per.dtype == torch.float64 and torch.allclose(mine, per, rtol=1e-12, atol=0), (nss_r, mine, per)
        assert torch.allclose(mo.nss(sal.view(B, 224, 224), fix_t, nss_r), whole, rtol=1e-12, atol=0)
        out["per_sample"][nss_r] = per
        out["batch"][nss_r] = whole
        print("nss_r", nss_r, "reference NSS per sample", per.tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nss_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
