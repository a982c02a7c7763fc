"""Debug: ET golden case -- GPU output vs golden output and the loss terms. (GPU box)"""
import os, sys, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as mo
from avdn_b200 import _lib
from avdn_b200.models.ET_haa import ET
ARGS = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                             num_input_actions=1, dropout_emb=0.0)
g = torch.load(os.path.join(ROOT, "tests/golden/model_golden.pt"), weights_only=False)["et"]
torch.manual_seed(1)
et = ET(ARGS).cuda().eval()
inp = dict(directions=g["directions"].cuda(), frames=g["frames"].cuda(), lenths=list(g["lenths"]),
           lang=g["lang"].cuda(), lang_cls=g["lang_cls"].cuda())
out, hs = et.forward_features(**inp)
print("out gpu\n", out.cpu(), "\nout golden\n", g["output"])
print("hs rel", ((hs.cpu() - g["h_sali"]).abs().max() / g["h_sali"].abs().max()).item())
B = 2
gt = torch.from_numpy(np.unpackbits(g["gt_sal_packed"])[:B * 224 * 224].reshape(B, 224, 224).astype(np.float64))
sal = torch.nn.functional.interpolate(hs.cpu().view(B, 1, 8, 8), size=(224, 224), mode="bilinear", align_corners=False)
for name, o in (("gpu out", out.cpu()), ("golden out", g["output"])):
    for i in range(B):
        l = mo.et_loss(o[i:i + 1], sal[i:i + 1], g["gt_xy"][i:i + 1], g["gt_alt"][i:i + 1], g["gt_prog"][i:i + 1], gt[i:i + 1], nss_w=0.1)
        print(name, i, float(l))
att = (gt * 255).to(torch.uint8).cuda()
loss = torch.zeros(1, dtype=torch.float64, device="cuda"); loss_i = torch.zeros(B, dtype=torch.float64, device="cuda")
d_out = torch.zeros(B, 4, device="cuda"); d_hs = torch.zeros(B, 64, device="cuda")
ptr = _lib.ptr
_lib.call("avdn_loss", ptr(out.detach()), ptr(hs.detach()), ptr(g["gt_xy"].cuda()), ptr(g["gt_alt"].cuda()),
          ptr(g["gt_prog"].cuda()), ptr(att), None, B, 0.1, 0, 0.2 / B, ptr(loss), ptr(loss_i), ptr(d_out), ptr(d_hs))
print("kernel loss_i", loss_i.cpu(), "total", loss.item(), "golden", g["loss"].item())
