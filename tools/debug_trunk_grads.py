"""Layer-wise backward parity of the Darknet trunk (teacher forced): for every conv block the
torch-fp32 autograd of THAT block is fed our input activation and our upstream gradient, so
each row shows one layer of bf16 error (dz of BN-backward, dw of wgrad, dx of dgrad).  The
last columns are the end-to-end parameter-gradient errors against a full fp32 autograd pass.
Usage (GPU box): python tools/debug_trunk_grads.py [N=4]"""
import os
import sys
import tempfile

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as mo  # cfg text + fp32 restatement
from avdn_b200 import _lib
from avdn_b200.models import dark_net as DN

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4


def rel2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
    f.write(mo.yolov3_trunk_cfg())
torch.manual_seed(0)
net = DN.Darknet(f.name, 224).cuda().train()
x = torch.randn(N, 3, 224, 224, device="cuda")
dy = torch.randn(N, 512, 7, 7, device="cuda")

# ---- full fp32 reference (autograd) on the same weights
sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
      for k, v in net.state_dict().items()}
yo = mo.darknet_forward(x, sd, mo.yolov3_trunk_cfg(), train=True)
yo.backward(dy)
sdb = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k)
       for k, v in net.state_dict().items()}
yb = mo.darknet_forward(x, sdb, mo.yolov3_trunk_cfg(), train=True, storage="bf16")
yb.backward(dy)
print("bf16-storage oracle vs fp32 oracle, forward rel2:", rel2(yb, yo))

# ---- ours, with the backward loop opened up for snapshots
y = net(x)
print("forward rel2 vs fp32:", rel2(y, yo), " vs bf16-storage oracle:", rel2(y, yb))
eng = list(net._engines.values())[0]
eng.build_bwd(net)
for L in eng.layers:
    L.dw.zero_(); L.dgamma.zero_(); L.dbeta.zero_()
call, ptr = _lib.call, _lib.ptr
last = eng.last
call("avdn_nchw_f32_to_nhwc", ptr(dy), ptr(last.g), eng.N, last.Hout * last.Wout, last.Cout_p)
print("idx Cin Cout k s Hout |   dz     dgamma  dbeta  |  dw(layer) dx(layer) | dw(e2e) dgamma(e2e) | dw,dgamma vs bf16-oracle | bf16-oracle dw vs fp32")
for li in reversed(range(len(eng.layers))):
    L = eng.layers[li]
    conv, bn = net.module_list[L.idx][0], net.module_list[L.idx][1]
    g_in = L.g[..., :L.Cout].float().permute(0, 3, 1, 2).clone()
    src_g_before = None
    if not L.first:
        src_g_before = L.src.g.float().clone()
    dgam0, dbet0 = L.dgamma.clone(), L.dbeta.clone()
    DN._layer_backward(eng, L)
    # ---- teacher-forced reference of this block
    if L.first:
        xin = eng.x_in[..., :3].float().permute(0, 3, 1, 2)
    else:
        xin = L.src.a[..., :L.Cin].float().permute(0, 3, 1, 2)
    xin = xin.clone().requires_grad_(True)
    w = conv.weight.detach().clone().requires_grad_(True)
    gm = bn.weight.detach().clone().requires_grad_(True)
    bt = bn.bias.detach().clone().requires_grad_(True)
    z = F.conv2d(xin, w, None, stride=L.s, padding=(L.k - 1) // 2)
    z.retain_grad()
    a = F.leaky_relu(F.batch_norm(z, None, None, gm, bt, True, 0.1, 1e-5), 0.01)
    a.backward(g_in)
    dz_o = L.dz[..., :L.Cout].float().permute(0, 3, 1, 2)
    r_dz = rel2(dz_o, z.grad)
    r_dg = rel2(L.dgamma - dgam0, gm.grad)
    r_db = rel2(L.dbeta - dbet0, bt.grad)
    r_dw = rel2(L.dw, w.grad)
    if L.first:
        r_dx = float("nan")
    else:
        dx_o = (L.src.g.float() - (src_g_before if any(p.desc.core.accumulate for p in L.p_dgrad) else 0))
        r_dx = rel2(dx_o[..., :L.Cin].permute(0, 3, 1, 2), xin.grad)
    i = L.idx
    e_dw = rel2(L.dw, sd[f"module_list.{i}.conv_{i}.weight"].grad)
    e_dg = rel2(L.dgamma, sd[f"module_list.{i}.batch_norm_{i}.weight"].grad)
    print(f"{i:3d} {L.Cin:4d} {L.Cout:4d} {L.k} {L.s} {L.Hout:4d} | {r_dz:7.4f} {r_dg:7.4f} {r_db:7.4f} | "
          f"{r_dw:8.4f} {r_dx:8.4f} | {e_dw:7.4f} {e_dg:7.4f} | "
          f"{rel2(L.dw, sdb[f'module_list.{i}.conv_{i}.weight'].grad):7.4f} "
          f"{rel2(L.dgamma, sdb[f'module_list.{i}.batch_norm_{i}.weight'].grad):7.4f} | "
          f"{rel2(sdb[f'module_list.{i}.conv_{i}.weight'].grad, sd[f'module_list.{i}.conv_{i}.weight'].grad):7.4f}")
