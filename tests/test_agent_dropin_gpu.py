"""GPU: the reference's agent loop with its own signatures -- ``NavCMTAgent.rollout(train_ml, not_in_train, nss_w)``,
``.train(loader, n_epochs, feedback, nss_w_weighting)``, ``.test(loader, ...)`` over ``agent.env`` (an ``ANDHNavBatch``),
src/xview_et/agent.py:191-254,512-894.

The loss of one teacher-feedback training rollout is compared with the oracle's per-step autograd-free restatement
(cv2-exact attention maps, the oracle teacher at OUR poses, the oracle simulator on the teacher's actions, the oracle
BERT + ET on OUR trunk features) at 1e-2 relative; trajectory dicts, ``self.loss``, ``self.logs['IL_loss']`` and the
two-rollout ``train`` iteration are checked for the reference's bookkeeping."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import bert_oracle as bo
from oracle import model_oracle as mo
from oracle import teacher_oracle as to
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu

SIZE = 1024
BL, TR = np.array([40.0, -75.0]), np.array([40.02, -74.98])
LAT_RATIO = 0.02 / SIZE


class _Tok:
    """Stand-in for BertTokenizerFast (no vocabulary offline): words hash to ids; [CLS]=1 first, [SEP]=2 last, pad 0."""

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, texts, padding=True, return_tensors="pt"):
        rows = [[1] + [5 + (sum(ord(c) * (k + 1) for k, c in enumerate(w)) % (self.vocab - 5)) for w in t.split()] + [2]
                for t in texts]
        n = max(len(r) for r in rows)
        ids = torch.zeros(len(rows), n, dtype=torch.long)
        mask = torch.zeros(len(rows), n, dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r)
            mask[i, :len(r)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def _items(B, seed):
    rng = np.random.default_rng(seed)
    items = []
    for i in range(B):
        ctr = np.array([40.01, -74.99]) + rng.uniform(-0.004, 0.004, size=2)
        half = rng.uniform(0.0008, 0.0016)
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        sq = (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * half) @ R.T
        step = rng.uniform(-0.0015, 0.0015, size=2)
        path = [ctr + sq + k * step for k in range(int(rng.integers(2, 5)))]
        ang = round(mo.get_direction(np.mean(path[0], axis=0), (path[0][0] + path[0][1]) / 2)) % 360
        items.append(dict(map_name="m0", route_index=str(i), gps_botm_left=BL.copy(), gps_top_right=TR.copy(),
                          lng_ratio=LAT_RATIO, lat_ratio=LAT_RATIO, angle=float(ang),
                          gt_path_corners=[np.asarray(p) for p in path],
                          instructions="[ins] fly towards the %s building and stop" % ("red", "tall", "grey", "round")[i % 4],
                          pre_dialogs="[que] where should i go " if i % 2 else ""))
    return items


@pytest.fixture(scope="module")
def setup(built_lib):
    from transformers import BertConfig
    from avdn_b200.env import ANDHNavBatch
    from avdn_b200.models.bert import CustomBERTModel
    from avdn_b200.xview_et.agent import NavCMTAgent
    B, V = 4, 400
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name, darknet_weight_file=None,
                                 lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2, teacher_weight=1.0, max_action_len=4,
                                 no_dropout=True, train_val_on_full=False, vision_only=False, no_direction=False,
                                 optim="adamW")
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    agent.tokenizer = _Tok(V)
    agent.attach_lang_model(CustomBERTModel(BertConfig(num_hidden_layers=1, vocab_size=V, max_position_embeddings=64)))
    env = ANDHNavBatch(batch_size=B, device="cuda")
    tile = wo.synthetic_tile(seed=4, size=SIZE)
    att = wo.synthetic_attention_tile(seed=4, size=SIZE)
    env.map_batch["m0"], env.attention_map_batch["m0"] = tile, att
    env.batch = _items(B, 7)
    agent.env = env
    return agent, env, tile, att


def test_teacher_training_rollout_loss_vs_oracle(setup):
    agent, env, tile, att = setup
    B = env.batch_size
    for opt in agent.optimizers:
        opt.lr, opt.wd = 0.0, 0.0
        opt.zero_grad()
    agent.loss = 0
    agent.feedback = "teacher"
    agent.env_name = "train"
    n_log = len(agent.logs["IL_loss"])
    traj = agent.rollout(train_ml=0.2, nss_w=0.1)
    ours = float(agent.loss)
    lr = agent._last_rollout
    steps = lr["steps"]
    corners = lr["corners"].cpu().numpy()                 # [steps, B, 4, 2]
    ended = lr["ended"].cpu().numpy().astype(bool)
    # ---- the pose sequence is the oracle simulator driven by the oracle teacher ----
    items = env.batch
    c = np.stack([np.asarray(it["gt_path_corners"][0]) for it in items])
    d = np.array([it["angle"] for it in items], dtype=np.float64)
    e = np.zeros(B, dtype=bool)
    bounds = np.tile(np.concatenate([BL, TR]), (B, 1))
    tx, ta, tp = lr["tgt_xy"].cpu().numpy(), lr["tgt_alt"].cpu().numpy(), lr["tgt_prog"].cpu().numpy()
    for t in range(steps):
        assert np.allclose(corners[t], c, rtol=0, atol=1e-9), t
        out = np.zeros((B, 4), dtype=np.float32)
        for i in range(B):
            xy, alt, prog = to.teacher_action(c[i], items[i]["gt_path_corners"], e[i], feedback="teacher")
            assert np.allclose(tx[t, i], xy, atol=2e-5) and abs(ta[t, i] - alt) < 1e-5 and abs(tp[t, i] - prog) < 1e-5
            out[i] = (tx[t, i, 0], tx[t, i, 1], ta[t, i], tp[t, i])      # the simulator sees OUR targets
        c, d, e, *_ = mo.waypoint_step(out, c, bounds, d, e, 0.5, t == agent.args.max_action_len - 1)
        assert np.array_equal(ended[t], e), t
    # ---- loss: oracle BERT + ET per step on OUR trunk features, cv2-exact attention maps, OUR targets ----
    batch = lr["batch"]
    T = steps
    frames = agent._ctx[2]["frames"].detach().cpu().view(B, T, 512, 49)
    sd_b = {k: v.detach().cpu().clone() for k, v in agent.lang_model.state_dict().items() if "position_ids" not in k}
    sd_e = {k: v.detach().cpu().clone() for k, v in agent.vln_model.state_dict().items()}
    with torch.no_grad():
        seq, _, _ = bo.custom_bert_forward(sd_b, batch["input_ids"].cpu(), batch["attention_mask"].cpu())
        _, lin, _ = bo.custom_bert_forward(sd_b, batch["cls_input_ids"].cpu(), batch["cls_attention_mask"].cpu())
        dirs = batch["directions"].cpu()
        px = batch["corners_px"].cpu().numpy()
        total = 0.0
        for t in range(T):
            sal_gt = np.stack([wo.warp_fixed_point(att, wo.inverse_homography(px[i, t]))[:, :, 0] for i in range(B)])
            lt = [int(lr["lenths"][i][t]) for i in range(B)]
            out, sal, _ = mo.et_forward(sd_e, dirs[:, :t + 1], frames[:, :t + 1], lt, seq, lin)
            total = total + mo.step_loss(mo.et_loss(out, sal, batch["gt_xy"][:, t].cpu(), batch["gt_alt"][:, t].cpu(),
                                                    batch["gt_prog"][:, t].cpu(),
                                                    torch.from_numpy(sal_gt.astype(np.float64) / 255), 0.1), 0.2, B)
    ref = float(total)
    assert abs(ours - ref) <= 1e-2 * abs(ref), (ours, ref)
    # ---- bookkeeping of the reference ----
    assert len(agent.logs["IL_loss"]) == n_log + 1 and abs(agent.il_losses()[-1] - ours) < 1e-9
    assert len(traj) == B
    for i, tr in enumerate(traj):
        assert tr["instr_id"] == "m0__" + str(i) and tr["num_dia"] == 1
        assert len(tr["actions"]) == len(tr["gt_actions"]) == len(tr["gt_progress"]) == len(tr["progress"])
        assert 1 <= len(tr["actions"]) <= steps
        assert len(tr["path_corners"]) in (len(tr["actions"]), len(tr["actions"]) + 1)
        assert np.array_equal(tr["path_corners"][0][0], np.asarray(items[i]["gt_path_corners"][0]))
    # gradients reached all three models
    for opt in agent.optimizers:
        assert float(opt.g.abs().sum()) > 0


def test_train_and_test_loops(setup):
    """``train(loader, 1, feedback='student')``: teacher rollout WITHOUT the NSS term + student rollout with it, one
    optimiser step; ``test(loader)``: one trajectory per episode that ``eval_metrics`` can score."""
    agent, env, tile, att = setup
    B = env.batch_size
    for opt in agent.optimizers:
        opt.lr, opt.wd = 1e-5, 0.0
    p0 = [opt.p.clone() for opt in agent.optimizers]
    n_log = len(agent.logs["IL_loss"])
    loader = [None]                                       # the reference's loader fills env as a side effect
    agent.train(loader, 1, feedback="student")
    assert len(agent.logs["IL_loss"]) == n_log + 2        # two rollouts
    assert all(not torch.equal(opt.p, q) for opt, q in zip(agent.optimizers, p0))
    assert all(torch.isfinite(opt.p).all() for opt in agent.optimizers)
    # the teacher half carries no NSS term: its logged loss equals a teacher rollout with nss_w = 0
    for opt in agent.optimizers:
        opt.lr = 0.0
        opt.zero_grad()
    agent.feedback, agent.loss = "teacher", 0
    agent.rollout(train_ml=0.2, nss_w=0)
    l0 = float(agent.loss)
    agent.loss = 0
    for opt in agent.optimizers:
        opt.zero_grad()
    agent.rollout(train_ml=0.2, nss_w=0.1)
    assert float(agent.loss) != l0                         # the NSS term is really switched by the argument
    # ---- test(): student feedback, eval mode ----
    res = agent.test(loader, env_name="val_seen", feedback="student")
    assert sorted(res) == ["m0__%d" % i for i in range(B)]
    for tr in res.values():
        assert len(tr["gt_progress"]) == len(tr["progress"]) >= 1 and len(tr["path_corners"]) >= 1
    avg, _ = env.eval_metrics(res)
    assert set(("sr", "spl", "gp", "iou")) <= set(avg) and np.isfinite(avg["gp"])
    # teacher feedback in validation logs the human-attention scores (agent.py:683-693)
    res_t = agent.test(loader, env_name="val_seen", feedback="teacher")
    assert any("nss" in tr for tr in res_t.values())
    avg_h, _ = env.eval_metrics(res_t, human_att_eval=True)
    assert set(("HA_precision", "HA_recall", "nss")) <= set(avg_h)
