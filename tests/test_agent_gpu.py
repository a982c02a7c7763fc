"""GPU: the agent slice -- fused loss kernel, waypoint discretisation, flat-arena
AdamW + clipping, and the fused training step -- against the oracle.

Tolerances: discretised integers bit-exact; loss kernel (fp32/fp64 arithmetic) 1e-4
relative; the bf16 end-to-end step loss 1e-2 relative.
"""
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


def _args(cfg_path):
    return types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=cfg_path,
                                 darknet_weight_file=None, lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2,
                                 no_dropout=True)     # parity mode: the oracle arithmetic has no dropout


def _loss_case(B, seed, nss_r=0):
    g = torch.Generator().manual_seed(seed)
    output = torch.randn(B, 4, generator=g)
    h_sali = torch.relu(torch.randn(B, 64, generator=g))
    gt_xy = torch.rand(B, 2, generator=g) * 2 - 1
    gt_alt, gt_prog = torch.rand(B, generator=g), torch.rand(B, generator=g)
    rng = np.random.default_rng(seed)
    att = (rng.random((B, 224, 224)) > 0.9).astype(np.uint8) * rng.integers(1, 256, size=(B, 224, 224)).astype(np.uint8)
    att[1] = 0                                    # a sample without human attention: NSS skipped
    jitter = torch.rand(B, generator=g) * 1e-5
    return output, h_sali, gt_xy, gt_alt, gt_prog, torch.from_numpy(att), jitter


@pytest.mark.parametrize("nss_r", [0, 1, -1])
def test_loss_kernel_value_and_gradients_vs_oracle(built_lib, nss_r):
    from avdn_b200 import _lib
    import torch.nn.functional as F
    B = 5
    output, h_sali, gt_xy, gt_alt, gt_prog, att, jitter = _loss_case(B, 3, nss_r)
    o = output.clone().requires_grad_(True)
    h = h_sali.clone().requires_grad_(True)
    pred = F.interpolate(h.view(-1, 1, 8, 8), size=(224, 224), mode="bilinear", align_corners=False)
    gt_sal = att.double() / 255
    ref = mo.step_loss(mo.et_loss(o, pred, gt_xy, gt_alt, gt_prog, gt_sal, nss_w=0.1, nss_r=nss_r, jitter=jitter),
                       0.2, B)
    ref.backward()
    dev = "cuda"
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    loss_i = torch.zeros(B, dtype=torch.float64, device=dev)
    d_o = torch.zeros(B, 4, device=dev)
    d_h = torch.zeros(B, 64, device=dev)
    ptr = _lib.ptr
    t = [x.to(dev).contiguous() for x in (output, h_sali, gt_xy, gt_alt, gt_prog, att, jitter)]
    _lib.call("avdn_loss", *[ptr(x) for x in t], B, 0.1, nss_r, 0.2 / B, ptr(loss), ptr(loss_i), ptr(d_o), ptr(d_h))
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    assert (d_o.cpu() - o.grad).abs().max() <= 1e-4 * o.grad.abs().max()
    assert (d_h.cpu() - h.grad).abs().max() <= 1e-4 * h.grad.abs().max()


def test_upsample_matches_interpolate(built_lib):
    import torch.nn.functional as F
    from avdn_b200.models.ET_haa import _UpsampleFn
    h = torch.randn(3, 64, device="cuda", requires_grad=True)
    p = _UpsampleFn.apply(h)
    ref_h = h.detach().cpu().clone().requires_grad_(True)
    ref = F.interpolate(ref_h.view(-1, 1, 8, 8), size=(224, 224), mode="bilinear", align_corners=False)
    assert (p.cpu() - ref).abs().max() < 1e-5
    w = torch.randn(3, 1, 224, 224)
    (p * w.cuda()).sum().backward()
    (ref * w).sum().backward()
    assert (h.grad.cpu() - ref_h.grad).abs().max() <= 1e-4 * ref_h.grad.abs().max()


def test_waypoint_postprocess_bit_exact(built_lib):
    from avdn_b200.xview_et.agent import NavCMTAgent
    g = torch.Generator().manual_seed(0)
    B = 4096
    out = torch.randn(B, 4, generator=g) * 1.5
    out[:8] = torch.tensor([[0., 0., 0.5, 0.5], [1., 0., 2., -1.], [0., 1., 0.5, 0.50001], [-1., -1., 1e-3, 0.5],
                            [3., -4., 0.25, 0.75], [1e-8, -1e-8, 0.125, 0.4999], [-0.5, 0., 0.9986, 1.0],
                            [0., -2., 0.0014, 0.0]])
    edge = torch.rand(B, generator=g, dtype=torch.float64) * 1e-3 + 1e-4
    ang, dist, alt, stop = mo.postprocess_waypoints(out.numpy(), edge.numpy(), 0.5)
    res = NavCMTAgent.postprocess_waypoints(None, out.cuda(), edge.cuda(), 0.5)
    assert np.array_equal(res["angle"].cpu().numpy().astype(np.int64), ang)
    assert np.array_equal(res["altitude"].cpu().numpy().astype(np.int64), alt)
    assert np.array_equal(res["stop"].cpu().numpy().astype(bool), stop)
    assert np.array_equal(res["dist"].cpu().numpy(), dist)            # float64, same operation order


def test_fused_adamw_matches_torch_adamw_with_clipping(built_lib):
    from avdn_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(33, 7), (129,), (64, 64), (5,)]
    ours = {f"p{i}": torch.nn.Parameter(torch.randn(s, device="cuda")) for i, s in enumerate(shapes)}
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours.values()]
    opt = FusedAdamW(ours, lr=1e-2, max_norm=2.0)
    ropt = torch.optim.AdamW(ref, lr=1e-2)
    for step in range(4):
        grads = [torch.randn(s, device="cuda") * (10.0 if step % 2 == 0 else 0.01) for s in shapes]
        opt.zero_grad()
        for (n, _), g in zip(ours.items(), grads):
            opt.grads[n].add_(g)
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 2.0)
        opt.step()
        ropt.step()
        for (n, p), r in zip(ours.items(), ref):
            assert torch.allclose(p.data, r.data, rtol=2e-5, atol=2e-6), (step, n, (p.data - r.data).abs().max())


def test_move_view_corners_matches_reference_formula(built_lib):
    """Host float64 restatement (agent.py:285-384): heading 0 square, zoom/rotate/move."""
    from avdn_b200.xview_et.agent import NavCMTAgent, get_direction
    c = np.array([[0.5010, 0.5000], [0.5010, 0.5010], [0.5000, 0.5010], [0.5000, 0.5000]])
    assert round(get_direction(np.mean(c, 0), (c[0] + c[1]) / 2)) % 360 == 0
    edge_m = np.linalg.norm(c[1] - c[0]) * 11.13 * 1e4
    new, heading = NavCMTAgent.move_view_corners(None, c, 90, 0.0005, edge_m, (0.0, 0.0), (1.0, 1.0))
    assert heading == 90
    # same size (altitude == current edge), rotated by 90 deg clockwise, moved 0.0005 along the new heading
    assert abs(np.linalg.norm(new[1] - new[0]) - np.linalg.norm(c[1] - c[0])) < 1e-9
    ctr = np.mean(new, 0) - np.mean(c, 0)
    assert abs(np.linalg.norm(ctr) - 0.0005) < 1e-8
    # leaving the map: unchanged corners, unchanged heading
    same, h2 = NavCMTAgent.move_view_corners(None, c, 0, 0.0, edge_m * 2000, (0.0, 0.0), (1.0, 1.0))
    assert np.array_equal(same, c) and h2 == 0


@pytest.fixture(scope="module")
def agent(built_lib):
    from avdn_b200.xview_et.agent import NavCMTAgent
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    torch.manual_seed(0)
    a = NavCMTAgent(_args(f.name), device="cuda")
    # conditioned synthetic weights: the reference's default init with the BN gamma of every
    # residual branch scaled by 0.1 (perturbation gain ~7x instead of ~100x through the 57
    # blocks -- DESIGN.md, conditioning), so that end-to-end comparisons carry information
    defs = mo.parse_cfg_text(mo.yolov3_trunk_cfg())[1:]
    with torch.no_grad():
        for i, d in enumerate(defs):
            if d["type"] == "shortcut":
                a.vision_model.module_list[i - 1][1].weight.mul_(0.1)
    tile = wo.synthetic_tile(seed=2, size=1024)
    att = wo.synthetic_attention_tile(seed=2, size=1024)
    a.renderer.add_map("m", tile, att)
    a._tile, a._att = tile, att
    return a


def _small_batch(B, T, L, seed):
    g = torch.Generator().manual_seed(seed)
    corners = wo.synthetic_pose_corners(B * T, seed=seed, size=3000, edge_frac=0.0).astype(np.float64)
    corners = np.rint((corners - 1500) * 0.25 + 512).astype(np.int32).reshape(B, T, 4, 2)
    deg = torch.randint(0, 360, (B, T), generator=g).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    return dict(corners_px=torch.from_numpy(corners), lang=torch.randn(B, L, 768, generator=g),
                lang_cls=torch.relu(torch.randn(B, 49, generator=g)), directions=dirs,
                gt_xy=torch.rand(B, 2, generator=g) * 2 - 1, gt_alt=torch.rand(B, generator=g),
                gt_prog=torch.rand(B, generator=g), lenths=[T, T - 1][:B] + [T] * max(0, B - 2))


def test_train_step_loss_matches_oracle_and_learns(agent):
    B, T, L = 2, 2, 24
    hb = _small_batch(B, T, L, 11)
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    # ---- oracle on the same weights: cv2-exact views -> fp32 Darknet (train BN) -> ET -> loss
    sd_t = {k: v.detach().cpu().clone() for k, v in agent.vision_model.state_dict().items()}
    sd_e = {k: v.detach().cpu().clone() for k, v in agent.vln_model.state_dict().items()}
    views, atts = [], []
    for b in range(B):
        for t in range(T):
            Mi = wo.inverse_homography(hb["corners_px"][b, t].numpy())
            views.append(wo.warp_fixed_point(agent._tile, Mi))
            if t == T - 1:
                atts.append(wo.warp_fixed_point(agent._att, Mi)[:, :, 0])
    x = torch.from_numpy(wo.normalise_views(np.stack(views)))
    feats = mo.darknet_forward(x, sd_t, mo.yolov3_trunk_cfg(), train=True).view(B, T, 512, 49)
    out, sal, _ = mo.et_forward(sd_e, hb["directions"], feats, hb["lenths"], hb["lang"], hb["lang_cls"])
    gt_sal = torch.from_numpy(np.stack(atts).astype(np.float64) / 255)
    ref = float(mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"], hb["gt_alt"], hb["gt_prog"], gt_sal, 0.1), 0.2, B))
    p0 = agent.et_optimizer.p.clone()
    t0 = agent.vision_model_optimizer.p.clone()
    first = agent.train_step(batch, sync_loss=True)
    assert abs(first - ref) <= 1e-2 * abs(ref), (first, ref)
    assert agent.launches > 0
    assert not torch.equal(agent.et_optimizer.p, p0) and not torch.equal(agent.vision_model_optimizer.p, t0)
    assert torch.isfinite(agent.et_optimizer.p).all() and torch.isfinite(agent.vision_model_optimizer.p).all()
    # the same batch again and again: the loss must go down (lr raised for the test)
    for opt in agent.optimizers:
        opt.lr = 1e-4
    losses = [agent.train_step(batch, sync_loss=True) for _ in range(12)]
    assert losses[-1] < first, (first, losses)


def test_train_step_loss_at_config1_shape_vs_oracle(agent):
    """The fused ``train_step`` at BASELINE configs[0]'s shape (B = 4 episodes, T = 10 views each, L = 250 tokens,
    ragged ``lenths``) against the oracle pipeline on the same weights: cv2-exact views of all 40 poses -> fp32
    Darknet with train-mode BatchNorm over the 40 views -> ET -> loss incl. NSS.  Step loss within 1e-2 relative
    (bf16 path, north_star tolerance; conditioned synthetic trunk weights, see the ``agent`` fixture)."""
    B, T, L = 4, 10, 250
    hb = _small_batch(B, T, L, 31)
    hb["lenths"] = [10, 4, 7, 1]
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    sd_t = {k: v.detach().cpu().clone() for k, v in agent.vision_model.state_dict().items()}
    sd_e = {k: v.detach().cpu().clone() for k, v in agent.vln_model.state_dict().items()}
    views, atts = [], []
    for b in range(B):
        for t in range(T):
            Mi = wo.inverse_homography(hb["corners_px"][b, t].numpy())
            views.append(wo.warp_fixed_point(agent._tile, Mi))
            if t == T - 1:
                atts.append(wo.warp_fixed_point(agent._att, Mi)[:, :, 0])
    x = torch.from_numpy(wo.normalise_views(np.stack(views)))
    with torch.no_grad():
        feats = mo.darknet_forward(x, sd_t, mo.yolov3_trunk_cfg(), train=True).view(B, T, 512, 49)
        out, sal, _ = mo.et_forward(sd_e, hb["directions"], feats, hb["lenths"], hb["lang"], hb["lang_cls"])
        gt_sal = torch.from_numpy(np.stack(atts).astype(np.float64) / 255)
        ref = float(mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"], hb["gt_alt"], hb["gt_prog"], gt_sal, 0.1), 0.2, B))
    for opt in agent.optimizers:
        opt.lr = 0.0
        opt.wd = 0.0
    ours = agent.train_step(batch, sync_loss=True)
    assert abs(ours - ref) <= 1e-2 * abs(ref), (ours, ref)
    # and the forward-only entry point (trunk in eval mode is a different function: compare train_step's own output)
    o_dev = agent._ctx[1].output if hasattr(agent._ctx[1], "output") else None
    if o_dev is not None:
        assert (o_dev.float().cpu() - out).abs().max() <= 1e-2 * out.abs().max() + 5e-3


def test_step_gradients_vs_oracle(agent):
    """Gradients the fused step leaves in the arenas vs autograd of the oracle pipeline, run twice:
    float32 end to end (the reference arithmetic) and with the trunk's tensors stored in bf16
    (``storage="bf16"``).  We must be as close to the fp32 reference as bf16 storage allows."""
    B, T, L = 2, 2, 24
    hb = _small_batch(B, T, L, 12)
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in hb.items()}
    views, atts = [], []
    for b in range(B):
        for t in range(T):
            Mi = wo.inverse_homography(hb["corners_px"][b, t].numpy())
            views.append(wo.warp_fixed_point(agent._tile, Mi))
            if t == T - 1:
                atts.append(wo.warp_fixed_point(agent._att, Mi)[:, :, 0])
    x = torch.from_numpy(wo.normalise_views(np.stack(views)))
    gt_sal = torch.from_numpy(np.stack(atts).astype(np.float64) / 255)
    refs = []
    for storage in (None, "bf16"):
        sd_t = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
                for k, v in agent.vision_model.state_dict().items()}
        sd_e = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point())
                for k, v in agent.vln_model.state_dict().items()}
        feats = mo.darknet_forward(x, sd_t, mo.yolov3_trunk_cfg(), train=True, storage=storage).view(B, T, 512, 49)
        out, sal, _ = mo.et_forward(sd_e, hb["directions"], feats, hb["lenths"], hb["lang"], hb["lang_cls"])
        loss = mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"], hb["gt_alt"], hb["gt_prog"], gt_sal, 0.1), 0.2, B)
        loss.backward()
        refs.append((float(loss.detach()), sd_t, sd_e))
    for opt in agent.optimizers:
        opt.lr = 0.0                               # keep the weights: we only look at the gradients
        opt.wd = 0.0
    ours = agent.train_step(batch, sync_loss=True)
    assert abs(ours - refs[0][0]) <= 1e-2 * abs(refs[0][0]), (ours, refs[0][0])

    def rel2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

    report = {}
    for n in ("fc2.weight", "encoder_vl.enc_transformer.layers.0.self_attn.in_proj_weight",
              "encoder_vl.enc_transformer.layers.1.linear2.weight", "decoder_2_action_full.0.weight", "fc.0.weight"):
        e = rel2(agent.et_optimizer.grads[n], refs[0][2][n].grad)
        e16 = rel2(refs[1][2][n].grad, refs[0][2][n].grad)
        report[n] = (e, e16)
        # + the ET's own bf16 GEMM rounding (checked at 5e-2 on exact frames in test_et_gpu); at B=2, T=2
        # the attention-projection gradients are small differences of bf16 products
        assert e <= 1.5 * e16 + 0.15, (n, e, e16)
    for i in (79, 78, 62, 41, 12, 1, 0):
        n = f"module_list.{i}.conv_{i}.weight"
        e = rel2(agent.vision_model_optimizer.grads[n], refs[0][1][n].grad)
        e16 = rel2(refs[1][1][n].grad, refs[0][1][n].grad)
        report[n] = (e, e16)
        assert e <= 1.5 * e16 + 8e-2, (n, e, e16)
    print({k: (round(a, 4), round(b, 4)) for k, (a, b) in report.items()})


@pytest.mark.parametrize("nss_r", [0, 1, -1])
def test_loss_kernel_nss_vs_reference_agent_golden(built_lib, golden_dir, nss_r):
    """The NSS term of ``avdn_loss`` against the REFERENCE's ``NavCMTAgent.NSS`` values
    (tests/golden/nss_golden.pt, written by make_nss_golden.py from src/xview_et/agent.py:256-270): with zero
    waypoint error the per-sample loss is ``scale * nss_w * NSS_i``."""
    import os
    from avdn_b200 import _lib
    g = torch.load(os.path.join(golden_dir, "nss_golden.pt"), weights_only=False)
    B = g["h_sali"].shape[0]
    att = torch.from_numpy(np.unpackbits(g["fix_packed"])[:B * 224 * 224].reshape(B, 224, 224)) * 255
    dev = "cuda"
    # output == targets (up to the angular term's own convention): xy on the unit axis, alt / prog equal
    output = torch.tensor([[0.0, 1.0, 0.5, 0.25]] * B)
    gt_xy, gt_alt, gt_prog = output[:, :2].clone(), output[:, 2].clone(), output[:, 3].clone()
    jitter = torch.zeros(B)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    loss_i = torch.zeros(B, dtype=torch.float64, device=dev)
    d_o = torch.zeros(B, 4, device=dev)
    d_h = torch.zeros(B, 64, device=dev)
    t = [x.to(dev).contiguous() for x in (output, g["h_sali"], gt_xy, gt_alt, gt_prog, att.to(torch.uint8), jitter)]
    nss_w, scale = 0.1, 0.2 / B
    _lib.call("avdn_loss", *[_lib.ptr(x) for x in t], B, nss_w, nss_r, scale, _lib.ptr(loss), _lib.ptr(loss_i),
              _lib.ptr(d_o), _lib.ptr(d_h))
    ref_i = g["per_sample"][nss_r] * nss_w
    got_i = loss_i.cpu()
    # loss_i may or may not carry the step scale: accept the kernel's convention, pin the NSS values
    ratio = (got_i / ref_i)
    k = ratio[0].item()
    assert abs(k - 1.0) < 1e-4 or abs(k - scale) < 1e-4 * scale, k
    assert torch.allclose(got_i, ref_i * k, rtol=1e-4, atol=1e-9), (got_i, ref_i * k)
    assert abs(loss.item() - (ref_i.sum() * scale).item()) <= 1e-4 * abs((ref_i.sum() * scale).item())
