"""GPU: the HAA-Transformer's inference path -- ``NavCMTAgent.rollout_greedy`` of the ET agent (student
feedback loop of src/xview_et/agent.py:580-760) against the oracle pipeline.

Step 0 end to end (cv2-exact views -> fp32 trunk in eval mode -> ET); every later step teacher-forced:
the oracle ET is fed OUR feature / heading history and the reference's ``lenths`` bookkeeping, the oracle
simulator OUR network outputs (the discretisation may legitimately flip between an fp32 and a bf16 trunk,
the history handling and the update rule may not)."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-6)


def _poses(B, seed):
    rng = np.random.default_rng(seed)
    bl, tr = np.array([40.0, -75.0]), np.array([40.02, -74.98])
    corners = np.zeros((B, 4, 2))
    dirs = np.zeros(B)
    for i in range(B):
        ctr = np.array([40.01, -74.99]) + rng.uniform(-0.0085, 0.0085, size=2)
        half = rng.uniform(0.0004, 0.0018)
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        corners[i] = ctr + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * half) @ R.T
        dirs[i] = round(mo.get_direction(np.mean(corners[i], axis=0), (corners[i][0] + corners[i][1]) / 2)) % 360
    return corners, dirs, np.tile(np.concatenate([bl, tr]), (B, 1))


@pytest.fixture(scope="module")
def agent(built_lib):
    from avdn_b200.xview_et.agent import NavCMTAgent
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name,
                                 darknet_weight_file=None, lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2, max_action_len=5)
    torch.manual_seed(0)
    a = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    tile = wo.synthetic_tile(seed=2, size=1024)
    a.renderer.add_map("m", tile, None)
    a._tile = tile
    return a


@pytest.mark.parametrize("thr_mode,incremental", [("reference", True), ("median", True), ("median", False)])
def test_et_greedy_rollout_vs_oracle_pipeline(agent, thr_mode, incremental):
    B, T, L = 4, 5, 12
    size = 1024
    lat_ratio = 0.02 / size
    geo = np.tile(np.array([40.0, -75.0, 40.02, -74.98, lat_ratio]), (B, 1))
    corners, dirs, bounds = _poses(B, 3)
    g = torch.Generator().manual_seed(2)
    lang = torch.randn(B, L, 768, generator=g)
    cls = torch.relu(torch.randn(B, 49, generator=g))
    batch = dict(corners_gps=torch.from_numpy(corners).cuda(), directions=torch.from_numpy(dirs).cuda(),
                 geo=torch.from_numpy(geo).cuda(), tile_idx=None, lang=lang.cuda(), lang_cls=cls.cuda())
    thr = 0.5
    if thr_mode == "median":                       # make some (not all) episodes stop after the first step
        probe = agent.rollout_greedy(batch, max_action_len=T, stop_threshold=1e9, incremental=incremental)
        thr = float(probe["output"][0, :, 3].clamp(0, 1).median().item())
    res = agent.rollout_greedy(batch, max_action_len=T, stop_threshold=thr, incremental=incremental)
    torch.cuda.synchronize()
    steps = res["steps"]
    out = res["output"].cpu()
    ch = res["corners"].cpu().numpy()
    dh = res["directions"].cpu().numpy()
    eh = res["ended"].cpu().numpy().astype(bool)
    assert np.array_equal(ch[0], corners) and np.array_equal(dh[0], dirs)
    sd_t = {k: v.detach().cpu() for k, v in agent.vision_model.state_dict().items()}
    sd_e = {k: v.detach().cpu() for k, v in agent.vln_model.state_dict().items()}
    # ---- step 0 through the oracle pipeline ----
    px = np.zeros((B, 4, 2), dtype=np.int32)
    for i in range(B):
        for k in range(4):
            lat, lng = corners[i, k]
            px[i, k] = (int(round((lng - geo[i, 1]) / lat_ratio)), int(round((geo[i, 2] - lat) / lat_ratio)))
    views = np.stack([wo.warp_fixed_point(agent._tile, wo.inverse_homography(px[i])) for i in range(B)])
    x = torch.from_numpy(wo.normalise_views(views))
    d0 = torch.from_numpy(dirs).float()
    dirs0 = torch.stack([torch.sin(d0 / 180 * 3.14159), torch.cos(d0 / 180 * 3.14159)], -1).view(B, 1, 2)
    with torch.no_grad():
        feats = mo.darknet_forward(x, sd_t, mo.yolov3_trunk_cfg(), train=False).view(B, 1, 512, 49)
        o0, _, _ = mo.et_forward(sd_e, dirs0, feats, [1] * B, lang, cls)
    assert _rel(out[0], o0) < 2e-2, _rel(out[0], o0)
    # ---- later steps, teacher-forced on OUR history ----
    bf = agent._bufs[("rollout", B, T)]
    fh = bf["frames_hist"].cpu()                   # [T,B,512,49]
    dhist = bf["dirs_hist"].cpu()                  # [T,B,2]
    ended = np.zeros(B, dtype=bool)
    lens = [0] * B
    for t in range(steps):
        for i in range(B):
            if not ended[i]:
                lens[i] += 1                        # agent.py:617-619
        rad = torch.from_numpy(dh[t]).float() / 180 * 3.14159
        assert torch.allclose(dhist[t], torch.stack([torch.sin(rad), torch.cos(rad)], -1), atol=2e-6)
        with torch.no_grad():
            ot, _, _ = mo.et_forward(sd_e, dhist[:t + 1].permute(1, 0, 2), fh[:t + 1].permute(1, 0, 2, 3), list(lens),
                                     lang, cls)
        # the full recompute reproduces the reference for every sample; the incremental path for every sample
        # that is still alive (an ended sample's output is never used: agent.py:736-746 `continue`s)
        alive = torch.from_numpy(~ended) if incremental else torch.ones(B, dtype=torch.bool)
        assert alive.any()
        assert _rel(out[t][alive], ot[alive]) < 1e-2, (t, _rel(out[t][alive], ot[alive]))
        nc, nd, ended, ang, alt, dist = mo.waypoint_step(out[t].numpy(), ch[t], bounds, dh[t], ended, thr, t == T - 1)
        assert np.array_equal(res["angle"][t].cpu().numpy().astype(np.int64), ang), t
        assert np.array_equal(res["altitude"][t].cpu().numpy().astype(np.int64), alt), t
        assert np.array_equal(eh[t], ended), t
        assert np.array_equal(dh[t + 1], nd), t
        np.testing.assert_allclose(ch[t + 1], nc, rtol=1e-12, atol=0)
    assert ended.all() and eh[T - 1].all()
    if thr_mode == "median":
        assert 0 < eh[0].sum() < B                  # ragged lengths were exercised
    traj = agent.trajectories(res)
    assert len(traj) == B and all(len(p) >= 1 for p in traj)


def test_incremental_rollout_matches_full_recompute(agent):
    """The default rollout computes two new rows per step against cached keys / values; re-running the encoder
    over the whole history every step (what the reference does) must give the same network outputs for every
    sample that is still alive, and (here: no early stop) the same discretised trajectory."""
    B, T, L = 6, 5, 20
    size = 1024
    lat_ratio = 0.02 / size
    geo = np.tile(np.array([40.0, -75.0, 40.02, -74.98, lat_ratio]), (B, 1))
    corners, dirs, _ = _poses(B, 11)
    g = torch.Generator().manual_seed(4)
    batch = dict(corners_gps=torch.from_numpy(corners).cuda(), directions=torch.from_numpy(dirs).cuda(),
                 geo=torch.from_numpy(geo).cuda(), tile_idx=None, lang=torch.randn(B, L, 768, generator=g).cuda(),
                 lang_cls=torch.relu(torch.randn(B, 49, generator=g)).cuda())
    inc = {k: (v.clone() if torch.is_tensor(v) else v)
           for k, v in agent.rollout_greedy(batch, max_action_len=T, stop_threshold=2.0, incremental=True).items()}
    # teacher-force the full recompute on the incremental run's history: same poses -> same frames
    bf = agent._bufs[("rollout", B, T)]
    fh, dhist = bf["frames_hist"].clone(), bf["dirs_hist"].clone()
    et = agent.vln_model
    pe = et.encoder_vl.enc_pos.pe[0]
    for t in range(T):
        eng = et.engine(B, L, t + 1, "cuda")
        eng.set_dropout(0.0, 0.0, 0)
        o, _ = eng.forward(fh[:t + 1].permute(1, 0, 2, 3).contiguous().view(B * (t + 1), 512, 49), batch["lang"],
                           batch["lang_cls"], dhist[:t + 1].permute(1, 0, 2).contiguous(), [t + 1] * B, pe)
        assert _rel(inc["output"][t], o) < 1e-2, (t, _rel(inc["output"][t], o))
    full = agent.rollout_greedy(batch, max_action_len=T, stop_threshold=2.0, incremental=False)
    assert inc["steps"] == full["steps"] == T
    assert _rel(inc["output"][0], full["output"][0]) < 1e-6      # step 0 is the same computation


def test_agent_test_produces_scorable_trajectories(agent):
    """agent.test -> trajectory dicts -> ANDHNavBatch.eval_metrics (inference, supervision geometry and the
    evaluation glue end to end)."""
    from avdn_b200.env import ANDHNavBatch
    from oracle import teacher_oracle as to
    B, T, L = 4, 5, 12
    size = 1024
    lat_ratio = 0.02 / size
    geo = np.tile(np.array([40.0, -75.0, 40.02, -74.98, lat_ratio]), (B, 1))
    corners, dirs, _ = _poses(B, 3)
    g = torch.Generator().manual_seed(2)
    rng = np.random.default_rng(9)
    gts = []
    for i in range(B):
        path = [corners[i]]
        for _ in range(int(rng.integers(1, 4))):
            path.append(path[-1] + rng.uniform(-0.001, 0.001, 2))
        gts.append(np.stack(path))
    batch = dict(corners_gps=torch.from_numpy(corners).cuda(), directions=torch.from_numpy(dirs).cuda(),
                 geo=torch.from_numpy(geo).cuda(), tile_idx=None, lang=torch.randn(B, L, 768, generator=g).cuda(),
                 lang_cls=torch.relu(torch.randn(B, 49, generator=g)).cuda(),
                 instr_id=[f"ep{i}" for i in range(B)], gt_path_corners=gts, num_dia=[1, 2, 3, 1])
    results = agent.test([batch], env_name="val_seen", max_action_len=T)
    assert sorted(results) == [f"ep{i}" for i in range(B)]
    for i in range(B):
        tr = results[f"ep{i}"]
        n = len(tr["gt_progress"])
        assert 1 <= n <= T and len(tr["actions"]) == n and len(tr["progress"]) == n
        assert len(tr["path_corners"]) in (n, n + 1) and np.array_equal(tr["path_corners"][0][0], corners[i])
        for k in range(min(n, len(tr["path_corners"]))):
            ref = to.compute_iou(tr["path_corners"][k][0], gts[i][-1])
            assert abs(tr["gt_progress"][k] - ref) < 1e-5, (i, k)
    env = ANDHNavBatch.__new__(ANDHNavBatch)
    avg, metrics = env.eval_metrics(results)
    assert set(("sr", "spl", "gp", "iou", "lengths")) <= set(avg) and len(metrics["instr_id"]) == B
    assert all(np.isfinite(float(v)) for v in avg.values())


def test_student_batch_feeds_the_training_rollout(agent):
    """agent.py:245-251: the student half of an iteration -- poses from a greedy rollout, per-step targets from
    ``teacher_action`` (checked against the oracle's teacher), then ``train_rollout_step`` on that batch."""
    from oracle import teacher_oracle as to
    B, T, L = 4, 4, 12
    size = 1024
    geo = np.tile(np.array([40.0, -75.0, 40.02, -74.98, 0.02 / size]), (B, 1))
    corners, dirs, _ = _poses(B, 3)
    rng = np.random.default_rng(5)
    gts = []
    for i in range(B):
        path = [corners[i]]
        for _ in range(int(rng.integers(1, 4))):
            path.append(path[-1] + rng.uniform(-0.001, 0.001, 2))
        gts.append(np.stack(path))
    g = torch.Generator().manual_seed(2)
    batch = dict(corners_gps=torch.from_numpy(corners).cuda(), directions=torch.from_numpy(dirs).cuda(),
                 geo=torch.from_numpy(geo).cuda(), tile_idx=None, lang=torch.randn(B, L, 768, generator=g).cuda(),
                 lang_cls=torch.relu(torch.randn(B, 49, generator=g)).cuda())
    tb = agent.student_batch(batch, gts, max_action_len=T)
    Ts = tb["directions"].shape[1]
    assert tb["corners_px"].shape == (B, Ts, 4, 2) and tb["gt_xy"].shape == (B, Ts, 2)
    assert np.array_equal(np.asarray(tb["lenths"])[:, 0], np.ones(B, dtype=int))
    res_corners = agent._bufs[("rollout", B, T)]["corners_hist"].cpu().numpy()
    ended = agent._bufs[("rollout", B, T)]["ended_hist"].cpu().numpy().astype(bool)
    for t in range(Ts):
        eb = ended[t - 1] if t > 0 else np.zeros(B, dtype=bool)
        for i in range(B):
            r, a, p = to.teacher_action(res_corners[t, i], gts[i], bool(eb[i]), feedback="student")
            assert np.allclose(tb["gt_xy"][i, t].cpu().numpy(), r, atol=1e-5), (t, i)
            assert abs(float(tb["gt_alt"][i, t]) - a) < 1e-5 and abs(float(tb["gt_prog"][i, t]) - p) < 1e-5
    agent.args.no_dropout = True
    l0 = agent.train_rollout_step(tb, sync_loss=True)
    losses = [agent.train_rollout_step(tb, sync_loss=True) for _ in range(6)]
    assert np.isfinite([l0] + losses).all() and losses[-1] < l0, (l0, losses)


def _shallow_cfg():
    """Seven convolutions down to [512,7,7]: a trunk shallow enough that differences between two of our own runs
    (reduce order of the weight-gradient adds) are not amplified."""
    out = ["[net]", "channels=3", "height=224", ""]
    for f, k, st in ((32, 3, 1), (64, 3, 2), (128, 3, 2), (256, 3, 2), (512, 3, 2), (1024, 3, 2), (512, 1, 1)):
        out.extend(["[convolutional]", "batch_normalize=1", f"filters={f}", f"size={k}", f"stride={st}", "pad=1",
                    "activation=leaky", ""])
    return "\n".join(out)


def test_train_iteration_accumulates_both_rollouts(built_lib):
    """agent.py:225-251: the gradients of the teacher and the student rollout add before the one optimiser step."""
    from avdn_b200.xview_et.agent import NavCMTAgent
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(_shallow_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name,
                                 darknet_weight_file=None, lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2, no_dropout=True)
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    B, T, L = 2, 3, 8
    g = torch.Generator().manual_seed(8)

    def mk():
        deg = torch.randint(0, 360, (B, T), generator=g).float()
        images = torch.zeros(B * T, 224, 224, 4)
        images[..., :3] = torch.randn(B * T, 224, 224, 3, generator=g)
        return {k: v.cuda() for k, v in dict(
            directions=torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1),
            images=images.bfloat16(), lang=torch.randn(B, L, 768, generator=g),
            lang_cls=torch.relu(torch.randn(B, 49, generator=g)), gt_xy=torch.rand(B, T, 2, generator=g) * 2 - 1,
            gt_alt=torch.rand(B, T, generator=g), gt_prog=torch.rand(B, T, generator=g)).items()}
    a, b = mk(), mk()
    saved = [(o.lr, o.wd) for o in agent.optimizers]
    for o in agent.optimizers:
        o.lr, o.wd = 0.0, 0.0
    try:
        rel = lambda x, y: ((x - y).norm() / y.norm().clamp_min(1e-30)).item()
        la = float(agent.train_rollout_step(a).item())
        ga = [o.g.clone() for o in agent.optimizers]
        agent.train_rollout_step(a)
        noise = [rel(o.g, x) for o, x in zip(agent.optimizers, ga)]      # run-to-run (atomics)
        lb = float(agent.train_rollout_step(b).item())
        gb = [o.g.clone() for o in agent.optimizers]
        tot = agent.train_iteration(a, b, sync_loss=True)
        assert abs(tot - (la + lb)) <= 2e-3 * abs(la + lb)
        errs = [rel(o.g, x + y) for o, x, y in zip(agent.optimizers, ga, gb)]
        print("additivity", errs, "run-to-run", noise)
        for e, nz in zip(errs, noise):
            assert nz < 1e-3 and e < 2e-2, (errs, noise)                 # a dropped rollout would show as ~0.5-1
    finally:
        for o, (lr, wd) in zip(agent.optimizers, saved):
            o.lr, o.wd = lr, wd
