"""Run-to-run reproducibility of the trunk's forward/backward (fp32 reduce-add order is not fixed: tiny
differences are expected; a race would show up as an occasional large one).  GPU box."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as mo
from avdn_b200.models.dark_net import Darknet
g = torch.load(os.path.join(ROOT, "tests/golden/model_golden.pt"), weights_only=False)["darknet"]
with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
    f.write(g["cfg"])
net = Darknet(f.name, 64).cuda().train()
x, dy = g["x"].cuda(), g["dy"].cuda()
ref = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    net.load_state_dict(g["sd"])
    net.zero_grad()
    y = net(x)
    y.backward(dy)
    cur = {n: p.grad.clone() for n, p in net.named_parameters()}
    cur["y"] = y.detach().clone()
    if ref is None:
        ref = cur
        continue
    worst = max(((cur[n] - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-30)).item() for n in ref)
    wn = max(ref, key=lambda n: ((cur[n] - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-30)).item())
    print(f"rep {rep}: worst rel diff {worst:.3e} at {wn}; y diff {(cur['y']-ref['y']).abs().max().item():.3e}")
