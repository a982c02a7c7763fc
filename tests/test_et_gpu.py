"""GPU: the CUDA HAA-Transformer (ET) -- tcgen05 GEMMs (bf16) + warp-level kernels --
against the committed golden vectors of the REFERENCE modules and against the fp32
oracle on identical weights.

Tolerances (north_star): masks / indices bit-exact; logits and loss within 1e-2
relative (bf16 tensor-core path); gradients of the bf16 path at 5e-2 relative L2.
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu

ARGS = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                             num_input_actions=1, dropout_emb=0.0)


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def _rel2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def golden(built_lib, golden_dir):
    return torch.load(os.path.join(golden_dir, "model_golden.pt"), weights_only=False)["et"]


@pytest.fixture(scope="module")
def et(golden):
    from avdn_b200.models.ET_haa import ET
    torch.manual_seed(1)                       # the fixture's recipe: default init under seed 1
    m = ET(ARGS)
    sd = m.state_dict()
    for k, v in golden["sd"].items():          # the small tensors are stored: they pin the recipe
        assert torch.equal(sd[k], v), k
    return m.cuda().eval()


def _inputs(g):
    return dict(directions=g["directions"].cuda(), frames=g["frames"].cuda().requires_grad_(True),
                lenths=list(g["lenths"]), lang=g["lang"].cuda(), lang_cls=g["lang_cls"].cuda())


def test_state_dict_keys_and_unused(et, golden):
    from avdn_b200.models.ET_haa import ET
    names = [n for n, _ in et.named_parameters()]
    used = set(et.used_parameters())
    assert sorted(set(names) - used) == sorted(golden["unused"])


def test_masks_bit_exact(et, golden):
    from avdn_b200.models import model_util
    L, T = golden["lang"].shape[1], golden["frames"].shape[1]
    ma = model_util.generate_attention_mask(L, T, "cuda")
    assert torch.equal(ma.cpu(), golden["mask_attn"])
    mp = model_util.generate_pad_mask(golden["lenths"], L, "cuda")
    assert torch.equal(mp.cpu(), golden["mask_pad"])
    # a larger ragged case against the oracle restatement
    lens = [10, 1, 7, 3, 10, 5]
    assert torch.equal(model_util.generate_pad_mask(lens, 250, "cuda").cpu(), mo.mask_pad(lens, 250))
    assert torch.equal(model_util.generate_attention_mask(250, 10, "cuda").cpu(), mo.attention_mask(250, 10))


def test_forward_vs_reference_golden(et, golden):
    out, sal = et(**_inputs(golden))
    assert out.shape == (2, 4) and sal.shape == (2, 1, 224, 224)
    assert _rel(out, golden["output"]) < 1e-2, _rel(out, golden["output"])
    assert _rel(sal[:, :, ::16, ::16], golden["sal_sub"]) < 1e-2
    o2, hs = et.forward_features(**_inputs(golden))
    assert _rel(hs, golden["h_sali"]) < 1e-2


def test_loss_and_gradients_vs_reference_golden(et, golden):
    from avdn_b200 import _lib
    g = golden
    inp = _inputs(g)
    out, hs = et.forward_features(**inp)
    B = 2
    att = torch.from_numpy(np.unpackbits(g["gt_sal_packed"])[: B * 224 * 224].reshape(B, 224, 224) * 255)
    att = att.to(torch.uint8).cuda()
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    loss_i = torch.zeros(B, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(B, 4, device="cuda")
    d_hs = torch.zeros(B, 64, device="cuda")
    ptr = _lib.ptr
    # keep the device copies alive across the call: ptr() of a temporary would dangle
    o_d, h_d = out.detach().contiguous(), hs.detach().contiguous()
    xy_d, alt_d, prog_d = g["gt_xy"].cuda(), g["gt_alt"].cuda(), g["gt_prog"].cuda()
    _lib.call("avdn_loss", ptr(o_d), ptr(h_d), ptr(xy_d), ptr(alt_d), ptr(prog_d), ptr(att), None, B, 0.1, 0,
              0.2 / B, ptr(loss), ptr(loss_i), ptr(d_out), ptr(d_hs))
    assert abs(loss.item() - g["loss"].item()) <= 1e-2 * abs(g["loss"].item()), (loss.item(), g["loss"].item())
    et.zero_grad()
    torch.autograd.backward([out, hs], [d_out, d_hs])
    for n, ref in g["grads"].items():
        p = dict(et.named_parameters())[n]
        assert p.grad is not None, n
        assert _rel2(p.grad, ref) < 5e-2, (n, _rel2(p.grad, ref))
    assert _rel2(inp["frames"].grad, g["frames_grad"]) < 5e-2
    for n in g["unused"]:
        assert dict(et.named_parameters())[n].grad is None


def test_all_gradients_vs_oracle_ragged(et):
    """B=3, L=40, T=5, ragged lengths: every used parameter's gradient against the fp32 oracle."""
    torch.manual_seed(5)
    B, L, T = 3, 40, 5
    lens = [5, 2, 4]
    lang = torch.randn(B, L, 768)
    lang_cls = torch.relu(torch.randn(B, 49))
    frames = torch.randn(B, T, 512, 49) * 0.5
    deg = torch.randint(0, 360, (B, T)).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    w_out = torch.randn(B, 4)
    w_hs = torch.randn(B, 64)
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in et.state_dict().items()}
    fr = frames.clone().requires_grad_(True)
    lg = lang.clone().requires_grad_(True)
    lc = lang_cls.clone().requires_grad_(True)
    oo, _, hs_o = mo.et_forward(sd, dirs, fr, lens, lg, lc)
    ((oo * w_out).sum() + (hs_o * w_hs).sum()).backward()
    f2 = frames.cuda().requires_grad_(True)
    l2 = lang.cuda().requires_grad_(True)
    c2 = lang_cls.cuda().requires_grad_(True)
    et.zero_grad()
    out, hs = et.forward_features(directions=dirs.cuda(), frames=f2, lenths=lens, lang=l2, lang_cls=c2)
    assert _rel(out, oo) < 1e-2 and _rel(hs, hs_o) < 1e-2
    ((out * w_out.cuda()).sum() + (hs * w_hs.cuda()).sum()).backward()
    worst = {}
    for n, p in et.used_parameters().items():
        r = _rel2(p.grad, sd[n].grad)
        worst[n] = r
        assert r < 5e-2, (n, r)
    assert _rel2(f2.grad, fr.grad) < 5e-2
    assert _rel2(l2.grad, lg.grad) < 5e-2
    assert _rel2(c2.grad, lc.grad) < 5e-2          # linear_cls feeds the trained BERT head in the reference


def test_config1_shape_forward_loss_and_all_gradients_vs_oracle(et):
    """BASELINE configs[0] shape -- B = 4, L = 250 dialog tokens, T = 10 views, S = 270 (padded to 288 rows inside the
    engine), ragged ``lenths`` with Tmax = 10: logits, saliency head, loss (fused ``avdn_loss`` incl. NSS) within 1e-2
    and EVERY used parameter's gradient (plus the gradients of frames / lang / lang_cls) within 5e-2 relative L2 of
    the fp32 oracle (oracle/model_oracle.py, pinned to the reference's ``ET`` by make_model_golden.py)."""
    from avdn_b200 import _lib
    torch.manual_seed(21)
    B, L, T = 4, 250, 10
    lens = [10, 3, 7, 1]
    lang = torch.randn(B, L, 768)
    lang_cls = torch.relu(torch.randn(B, 49))
    frames = torch.randn(B, T, 512, 49) * 0.5
    deg = torch.randint(0, 360, (B, T)).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    gt_xy = torch.rand(B, 2) * 2 - 1
    gt_xy = gt_xy / gt_xy.abs().amax(1, keepdim=True).clamp_min(1.0)
    gt_alt, gt_prog = torch.rand(B), torch.rand(B)
    rng = np.random.default_rng(21)
    att = np.zeros((B, 224, 224), dtype=np.uint8)
    for i in range(B - 1):                                   # the last sample has no attention: NSS skipped
        cy, cx, r = rng.integers(40, 180), rng.integers(40, 180), rng.integers(10, 50)
        yy, xx = np.ogrid[:224, :224]
        att[i][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
    gt_sal = torch.from_numpy(att.astype(np.float64) / 255)
    # ---- oracle ----
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in et.state_dict().items()}
    fr = frames.clone().requires_grad_(True)
    lg = lang.clone().requires_grad_(True)
    lc = lang_cls.clone().requires_grad_(True)
    oo, so, hs_o = mo.et_forward(sd, dirs, fr, lens, lg, lc)
    ref = mo.step_loss(mo.et_loss(oo, so, gt_xy, gt_alt, gt_prog, gt_sal, nss_w=0.1), 0.2, B)
    ref.backward()
    # ---- device ----
    f2 = frames.cuda().requires_grad_(True)
    l2 = lang.cuda().requires_grad_(True)
    c2 = lang_cls.cuda().requires_grad_(True)
    et.zero_grad()
    out, hs = et.forward_features(directions=dirs.cuda(), frames=f2, lenths=lens, lang=l2, lang_cls=c2)
    assert _rel(out, oo) < 1e-2, _rel(out, oo)
    assert _rel(hs, hs_o) < 1e-2, _rel(hs, hs_o)
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    loss_i = torch.zeros(B, dtype=torch.float64, device="cuda")
    d_out = torch.zeros(B, 4, device="cuda")
    d_hs = torch.zeros(B, 64, device="cuda")
    ptr = _lib.ptr
    o_d, h_d = out.detach().contiguous(), hs.detach().contiguous()
    xy_d, alt_d, prog_d, att_d = gt_xy.cuda(), gt_alt.cuda(), gt_prog.cuda(), torch.from_numpy(att).cuda()
    _lib.call("avdn_loss", ptr(o_d), ptr(h_d), ptr(xy_d), ptr(alt_d), ptr(prog_d), ptr(att_d), None, B, 0.1, 0,
              0.2 / B, ptr(loss), ptr(loss_i), ptr(d_out), ptr(d_hs))
    assert abs(loss.item() - ref.item()) <= 1e-2 * abs(ref.item()), (loss.item(), ref.item())
    torch.autograd.backward([out, hs], [d_out, d_hs])
    for n, p in et.used_parameters().items():
        r = _rel2(p.grad, sd[n].grad)
        assert r < 5e-2, (n, r)
    assert _rel2(f2.grad, fr.grad) < 5e-2
    assert _rel2(l2.grad, lg.grad) < 5e-2
    assert _rel2(c2.grad, lc.grad) < 5e-2


def test_encoder_vl_standalone(et):
    torch.manual_seed(7)
    B, L, T = 2, 20, 4
    lens = [4, 3]
    el, ef, ed = torch.randn(B, L, 768), torch.randn(B, T, 768), torch.randn(B, T, 768)
    sd = {k: v.detach().cpu() for k, v in et.state_dict().items()}
    ref, mp_ref = mo.encoder_vl_forward(el, ef, ed, lens, sd)
    out, mp = et.encoder_vl(el.cuda(), ef.cuda(), ed.cuda(), lens)
    assert torch.equal(mp.cpu(), mp_ref)
    # rows of padded steps are live queries in the reference too; compare everything
    assert _rel(out, ref) < 1e-2, _rel(out, ref)


def _site_mask(n, p, seed, site):
    from avdn_b200 import _lib
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.call("avdn_dropout_keep_scale", _lib.ptr(out), n, float(p), int(seed), int(site))
    return out.cpu()


def test_train_mode_dropout_vs_oracle_with_same_masks(et):
    """Train mode applies the reference's nn.Dropout sites (0.1 x 4 per encoder layer, 0.2 x 3 in the
    heads) with stateless hash masks.  The masks of the step are read back (avdn_dropout_keep_scale) and
    fed to the fp32 oracle as explicit masks: outputs and every gradient must then agree."""
    torch.manual_seed(9)
    B, L, T, H, S = 3, 40, 5, 12, 50
    Sp = 64
    lens = [5, 3, 4]
    lang = torch.randn(B, L, 768)
    lang_cls = torch.relu(torch.randn(B, 49))
    frames = torch.randn(B, T, 512, 49) * 0.5
    deg = torch.randint(0, 360, (B, T)).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    w_out, w_hs = torch.randn(B, 4), torch.randn(B, 64)
    f2 = frames.cuda().requires_grad_(True)
    l2 = lang.cuda().requires_grad_(True)
    et.train()
    try:
        et.zero_grad()
        out, hs = et.forward_features(directions=dirs.cuda(), frames=f2, lenths=lens, lang=l2,
                                      lang_cls=lang_cls.cuda())
        eng = et.engine(B, L, T, f2.device)
        p, ph, seed = eng.p_enc, eng.p_head, eng.seed
        assert p == pytest.approx(0.1) and ph == pytest.approx(0.2) and eng.Sp == Sp
        ((out * w_out.cuda()).sum() + (hs * w_hs.cuda()).sum()).backward()
        # a second train-mode forward draws different masks
        out_b, _ = et.forward_features(directions=dirs.cuda(), frames=f2.detach(), lenths=lens, lang=l2.detach(),
                                       lang_cls=lang_cls.cuda())
        assert eng.seed != seed and not torch.equal(out_b, out)
    finally:
        et.eval()
    layers = []
    for l in range(2):
        attn = _site_mask(B * H * S * Sp, p, seed, 4 * l).view(B, H, S, Sp)[..., :S].contiguous()
        layers.append(dict(attn=attn,
                           drop1=_site_mask(B * S * 768, p, seed, 4 * l + 1).view(B, S, 768),
                           ffn=_site_mask(B * S * 768, p, seed, 4 * l + 2).view(B, S, 768),
                           drop2=_site_mask(B * S * 768, p, seed, 4 * l + 3).view(B, S, 768)))
    drop = dict(layers=layers, h0=_site_mask(B * 256, ph, seed, 1000).view(B, 256),
                h1=_site_mask(B * 32, ph, seed, 1001).view(B, 32), fc=_site_mask(B * 64, ph, seed, 1002).view(B, 64))
    # the masks are Bernoulli(1-p) scaled by 1/(1-p)
    m = layers[0]["drop1"]
    u = torch.unique(m)
    assert len(u) == 2 and u[0].item() == 0.0 and abs(u[1].item() - 1 / 0.9) < 1e-6
    keep = (m > 0).float().mean().item()
    assert abs(keep - 0.9) < 4 * (0.9 * 0.1 / m.numel()) ** 0.5 + 1e-3, keep
    keep_a = (layers[1]["attn"] > 0).float().mean().item()
    assert abs(keep_a - 0.9) < 5e-3, keep_a
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in et.state_dict().items()}
    fr = frames.clone().requires_grad_(True)
    lg = lang.clone().requires_grad_(True)
    oo, _, hs_o = mo.et_forward(sd, dirs, fr, lens, lg, lang_cls, drop=drop)
    ((oo * w_out).sum() + (hs_o * w_hs).sum()).backward()
    assert _rel(out, oo) < 1e-2, _rel(out, oo)
    assert _rel(hs, hs_o) < 1e-2, _rel(hs, hs_o)
    for n, pp in et.used_parameters().items():
        r = _rel2(pp.grad, sd[n].grad)
        assert r < 5e-2, (n, r)
    assert _rel2(f2.grad, fr.grad) < 5e-2
    assert _rel2(l2.grad, lg.grad) < 5e-2
