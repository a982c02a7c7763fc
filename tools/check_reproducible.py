"""Run-to-run reproducibility of the train-mode trunk (forward features, statistics, gradients) on one GPU:
    python tools/check_reproducible.py [N]
Prints the relative differences between consecutive runs on identical inputs and weights."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__  # noqa: F401
from avdn_b200.models import dark_net as DN
from avdn_b200.utils import synthetic as syn

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
    f.write(syn.yolov3_trunk_cfg())
torch.manual_seed(0)
vm = DN.Darknet(f.name).cuda()
vm.train()
g = torch.Generator().manual_seed(8)
x = torch.zeros(N, 224, 224, 4)
x[..., :3] = torch.randn(N, 224, 224, 3, generator=g)
x = x.bfloat16().cuda()
eng = vm.engine(N, 224, 224, x.device)
out = torch.empty((N, 512, 7, 7), device="cuda")
dout = torch.randn(N, 512, 7, 7, generator=g).cuda() * 1e-3
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()
prev = None
for it in range(5):
    DN._trunk_forward(vm, eng, x, True, out=out)
    eng.build_bwd(vm)
    for L in eng.layers:
        L.dw.zero_(); L.dgamma.zero_(); L.dbeta.zero_()
    DN._trunk_backward(vm, eng, dout)
    torch.cuda.synchronize()
    cur = dict(features=out.clone(), mean=torch.cat([L.mean.flatten() for L in eng.layers]),
               rstd=torch.cat([L.rstd.flatten() for L in eng.layers]),
               dw=torch.cat([L.dw.flatten() for L in eng.layers]), dgamma=torch.cat([L.dgamma.flatten() for L in eng.layers]))
    if prev is not None:
        print(it, {k: "%.2e" % rel(cur[k], prev[k]) for k in cur})
    prev = {k: v.clone() for k, v in cur.items()}
