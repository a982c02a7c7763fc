"""GPU: the teacher-forced training rollout (src/xview_et/agent.py:580-760, 883-885): a loss at every step on the
growing history, summed, one backward.  ``train_rollout_step`` makes the reference's T encoder calls; the oracle
does the same with autograd on OUR trunk features (teacher-forced), so loss, the gradient that reaches the trunk
and the transformer's parameter gradients are compared."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu


def _rel2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("ragged", [False, True])
def test_rollout_step_matches_per_step_oracle(built_lib, ragged):
    from avdn_b200.xview_et.agent import NavCMTAgent
    B, T, L = 2, 3, 8
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name, darknet_weight_file=None,
                                 lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2, no_dropout=True)
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    for opt in agent.optimizers:
        opt.lr, opt.wd = 0.0, 0.0
    g = torch.Generator().manual_seed(4)
    deg = torch.randint(0, 360, (B, T), generator=g).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    images = torch.zeros(B * T, 224, 224, 4)
    images[..., :3] = torch.randn(B * T, 224, 224, 3, generator=g)
    att = np.zeros((B, T, 224, 224), np.uint8)
    att[0, :, 60:120, 80:160] = 255
    att[1, 1, 10:50, 150:200] = 255                      # steps without a fixation map carry no NSS term
    lens = [[1, 2, 3], [1, 2, 2]] if ragged else None     # sample 1 ended after its second step
    hb = dict(directions=dirs, images=images.bfloat16(), att=torch.from_numpy(att),
              lang=torch.randn(B, L, 768, generator=g), lang_cls=torch.relu(torch.randn(B, 49, generator=g)),
              gt_xy=torch.rand(B, T, 2, generator=g) * 2 - 1, gt_alt=torch.rand(B, T, generator=g),
              gt_prog=torch.rand(B, T, generator=g))
    batch = {k: v.cuda() for k, v in hb.items()}
    if lens is not None:
        batch["lenths"] = lens
    ours = agent.train_rollout_step(batch, sync_loss=True)
    bufs = agent._ctx[2]
    frames = bufs["frames"].detach().cpu().view(B, T, 512, 49).clone().requires_grad_(True)
    sd_e = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point())
            for k, v in agent.vln_model.state_dict().items()}
    total = 0
    for t in range(T):
        lt = [lens[i][t] for i in range(B)] if lens is not None else [t + 1] * B
        out, sal, _ = mo.et_forward(sd_e, dirs[:, :t + 1], frames[:, :t + 1], lt, hb["lang"], hb["lang_cls"])
        gt_sal = torch.from_numpy(att[:, t].astype(np.float64) / 255)
        total = total + mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"][:, t], hb["gt_alt"][:, t], hb["gt_prog"][:, t],
                                                gt_sal, 0.1), 0.2, B)
    total.backward()
    ref = float(total.detach())
    assert abs(ours - ref) <= 1e-2 * abs(ref), (ours, ref)
    d_frames = bufs["d_frames"].view(B, T, 512, 49)
    assert _rel2(d_frames, frames.grad) < 5e-2
    report = {}
    for n in ("encoder_vl.enc_transformer.layers.0.self_attn.in_proj_weight",
              "encoder_vl.enc_transformer.layers.1.linear2.weight", "direction_embedding.weight",
              "decoder_2_action_full.0.weight", "fc.0.weight"):
        if n in agent.et_optimizer.grads and sd_e[n].grad is not None:
            report[n] = _rel2(agent.et_optimizer.grads[n], sd_e[n].grad)
    print({k: round(v, 4) for k, v in report.items()})
    assert len(report) >= 3 and max(report.values()) < 0.1, report
    # the rollout step learns
    for opt in agent.optimizers:
        opt.lr = 1e-4
    losses = [agent.train_rollout_step(batch, sync_loss=True) for _ in range(8)]
    assert np.isfinite(losses).all() and losses[-1] < ours, (ours, losses)


def test_rollout_step_trains_the_language_encoder(built_lib):
    """The language encoder runs once per rollout (agent.py:519-538) and receives the gradients of every step."""
    from transformers import BertConfig
    from oracle import bert_oracle as bo
    from avdn_b200.models.bert import CustomBERTModel
    from avdn_b200.xview_et.agent import NavCMTAgent
    B, T, S, V = 2, 3, 12, 300
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name, darknet_weight_file=None,
                                 lr=1e-5, nss_w=0.0, nss_r=0, ml_weight=0.2, no_dropout=True)
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    lm = agent.attach_lang_model(CustomBERTModel(BertConfig(num_hidden_layers=2, vocab_size=V)))
    for opt in agent.optimizers:
        opt.lr, opt.wd = 0.0, 0.0
    g = torch.Generator().manual_seed(6)
    ids = torch.randint(0, V, (B, S), generator=g)
    mask = torch.ones(B, S, dtype=torch.long)
    mask[1, 8:] = 0
    deg = torch.randint(0, 360, (B, T), generator=g).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    images = torch.zeros(B * T, 224, 224, 4)
    images[..., :3] = torch.randn(B * T, 224, 224, 3, generator=g)
    hb = dict(input_ids=ids, attention_mask=mask, directions=dirs, images=images.bfloat16(),
              gt_xy=torch.rand(B, T, 2, generator=g) * 2 - 1, gt_alt=torch.rand(B, T, generator=g),
              gt_prog=torch.rand(B, T, generator=g))
    batch = {k: v.cuda() for k, v in hb.items()}
    ours = agent.train_rollout_step(batch, sync_loss=True)
    frames = agent._ctx[2]["frames"].detach().cpu().view(B, T, 512, 49)
    sd_b = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in lm.state_dict().items()
            if "position_ids" not in k}
    sd_e = {k: v.detach().cpu().clone() for k, v in agent.vln_model.state_dict().items()}
    seq, lin, _ = bo.custom_bert_forward(sd_b, ids, mask)
    total = 0
    zero_sal = torch.zeros(B, 224, 224, dtype=torch.float64)
    for t in range(T):
        out, sal, _ = mo.et_forward(sd_e, dirs[:, :t + 1], frames[:, :t + 1], [t + 1] * B, seq, lin)
        total = total + mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"][:, t], hb["gt_alt"][:, t], hb["gt_prog"][:, t],
                                                zero_sal, 0.0), 0.2, B)
    total.backward()
    ref = float(total.detach())
    assert abs(ours - ref) <= 1e-2 * abs(ref), (ours, ref)
    report = {}
    for n in ("bert.embeddings.word_embeddings.weight", "bert.encoder.layer.0.attention.self.query.weight",
              "bert.encoder.layer.1.output.dense.weight", "bert.encoder.layer.1.output.LayerNorm.weight"):
        report[n] = _rel2(agent.lang_optimizer.grads[n], sd_b[n].grad)
        assert sd_b[n].grad.norm() > 0, n
    print({k: round(v, 4) for k, v in report.items()})
    assert max(report.values()) < 0.1, report


def test_rollout_step_bn_per_step_is_the_reference_trunk_schedule(built_lib):
    """``bn_per_step=True``: one train-mode trunk pass per time step over that step's B views (src/xview_et/agent.py:593)
    -- batch statistics of B images, T running-statistics updates per rollout -- instead of one pass over B*T views.
    The features of step t must be bit-identical to a separate train-mode ``Darknet`` pass over the views of step t;
    with T = 1 the two schedules are the same computation (loss and gradients agree)."""
    from avdn_b200.models import dark_net as DN
    from avdn_b200.xview_et.agent import NavCMTAgent
    B, T, L = 3, 3, 8
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name, darknet_weight_file=None,
                                 lr=1e-5, nss_w=0.0, nss_r=0, ml_weight=0.2, no_dropout=True)
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    for opt in agent.optimizers:
        opt.lr, opt.wd = 0.0, 0.0
    g = torch.Generator().manual_seed(9)
    deg = torch.randint(0, 360, (B, T), generator=g).float()
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    images = torch.zeros(B * T, 224, 224, 4)
    images[..., :3] = torch.randn(B * T, 224, 224, 3, generator=g)
    hb = dict(directions=dirs, images=images.bfloat16(), lang=torch.randn(B, L, 768, generator=g),
              lang_cls=torch.relu(torch.randn(B, 49, generator=g)), gt_xy=torch.rand(B, T, 2, generator=g) * 2 - 1,
              gt_alt=torch.rand(B, T, generator=g), gt_prog=torch.rand(B, T, generator=g))
    batch = {k: v.cuda() for k, v in hb.items()}
    vm = agent.vision_model
    nb0 = int(vm.state_dict()["module_list.0.batch_norm_0.num_batches_tracked"])
    l_step = agent.train_rollout_step(batch, sync_loss=True, bn_per_step=True)
    assert np.isfinite(l_step)
    frames = agent._ctx[2]["frames"].view(B, T, 512, 49).clone()
    assert int(vm.state_dict()["module_list.0.batch_norm_0.num_batches_tracked"]) == nb0 + T
    x5 = batch["images"].view(B, T, 224, 224, 4)
    vm.train()
    for t in range(T):
        eng = vm.engine(B, 224, 224, "cuda", slot=100 + t)
        ref = DN._trunk_forward(vm, eng, x5[:, t].contiguous(), True)
        assert torch.equal(frames[:, t], ref.view(B, 512, 49)), t
    l_fused = agent.train_rollout_step(batch, sync_loss=True, bn_per_step=False)
    f_fused = agent._ctx[2]["frames"].view(B, T, 512, 49)
    assert not torch.equal(f_fused, frames)                 # statistics over B*T views are a different function
    # T = 1: the same computation under both schedules
    b1 = {k: (v[:, :1].contiguous() if k in ("directions", "gt_xy", "gt_alt", "gt_prog") else v) for k, v in batch.items()}
    b1["images"] = x5[:, 0].contiguous()
    la = agent.train_rollout_step(b1, sync_loss=True, bn_per_step=False)
    ga = agent.vision_model_optimizer.g.clone()
    lb = agent.train_rollout_step(b1, sync_loss=True, bn_per_step=True)
    gb = agent.vision_model_optimizer.g.clone()
    assert abs(la - lb) <= 1e-6 * abs(la)
    assert _rel2(gb, ga) < 1e-3, _rel2(gb, ga)
