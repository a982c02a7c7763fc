"""Per-layer CUDA-event timing of the Darknet trunk (fwd + bwd) at the bench shape.
Usage (GPU box): python tools/profile_layers.py [N=640] > gpurun_out/layers.txt"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as mo  # cfg text generator only
from avdn_b200 import _lib
from avdn_b200.models import dark_net as DN

N = int(sys.argv[1]) if len(sys.argv) > 1 else 640
with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
    f.write(mo.yolov3_trunk_cfg())
torch.manual_seed(0)
net = DN.Darknet(f.name, 224).cuda().train()
x = torch.randn(N, 224, 224, 4, device="cuda").bfloat16()
eng = net.engine(N, 224, 224, x.device)
dout = torch.randn(N, 512, 7, 7, device="cuda")
for it in range(3):
    if it == 2:
        _lib.PROFILE = []
    DN._trunk_forward(net, eng, x, True)
    DN._trunk_backward(net, eng, dout)
torch.cuda.synchronize()
rec = _lib.PROFILE
_lib.PROFILE = None
# map records back to layers by replaying the launch order
rows = []
it = iter(rec)
def take():
    n, e0, e1, fl, nb = next(it)
    return n, e0.elapsed_time(e1), fl
fw = {}
pack_all = take()            # one launch packs the weights of every block
for L in eng.layers:
    d = {"conv": take()}
    d["stats"] = take(); d["apply"] = take()
    fw[L.idx] = d
take()  # nhwc->nchw
take()  # nchw->nhwc
bw = {}
for L in reversed(eng.layers):
    d = {"bn_bwd": take()}
    if L.first and getattr(L, "recompute", False):
        d["wgrad"] = ("", 0.0, 0)        # avdn_conv0_bwd is one record: sums + weight gradient
    elif L.first:
        d["wgrad"] = take()
    else:
        d["wgrad"] = take()
        t, fl = 0.0, 0
        for _ in L.p_dgrad:
            n, ms, f2 = take(); t += ms; fl += f2
        d["dgrad"] = ("dgrad", t, fl)
    bw[L.idx] = d
print(f"# N={N}  per-layer ms and TFLOP/s (algorithmic);  act MB = z bytes")
print("idx  Cin Cout k s Hout  actMB |  fwd ms TF/s | stats apply | bnbwd | wgrad ms TF/s | dgrad ms TF/s")
tot = {}
for L in eng.layers:
    f_, b_ = fw[L.idx], bw[L.idx]
    act = L.R * L.Cout_p * 2 / 1e6
    tf = lambda r: (r[2] / (r[1] * 1e-3) / 1e12) if r[1] > 0 else 0
    dg = b_.get("dgrad", ("", 0.0, 0))
    print(f"{L.idx:3d} {L.Cin:4d} {L.Cout:4d} {L.k} {L.s} {L.Hout:4d} {act:7.0f} | {f_['conv'][1]:6.3f} {tf(f_['conv']):5.0f} | "
          f"{f_['stats'][1]:5.3f} {f_['apply'][1]:5.3f} | {b_['bn_bwd'][1]:6.3f} | {b_['wgrad'][1]:6.3f} {tf(b_['wgrad']):5.0f} | "
          f"{dg[1]:6.3f} {tf(dg):5.0f}")
    for k, v in list(f_.items()) + list(b_.items()):
        tot[k] = tot.get(k, 0.0) + v[1]
unpack_all = take()          # one launch folds every WGRAD output into the weight gradients
tot["pack"], tot["unpack"] = pack_all[1], unpack_all[1]
print("totals ms:", {k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2))
