"""GPU: BASELINE.json's full sizes through size-independent properties (the oracle cannot run these
sizes in seconds).  Config 2: the B=64 x T=10 x L=250 training step -- per-sample losses sum to the step
loss, permuting the episodes of the batch permutes the per-sample losses (BatchNorm statistics are
batch-wide, everything else is per episode), repeated steps are finite and learn.  Config 5: the
B=256 x 20-step greedy rollout -- episodes are independent in eval mode, so any shard of the batch
reproduces its part of the full-batch rollout exactly (what pose/episode sharding across GPUs relies on)."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from avdn_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu

SIZE = 3000


def _cfg_file():
    f = tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False)
    f.write(syn.yolov3_trunk_cfg())
    f.close()
    return f.name


def _train_batch(B, T, L, seed):
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    corners = syn.synthetic_pose_corners(B * T, seed=seed, size=SIZE, edge_frac=0.05).reshape(B, T, 4, 2)
    deg = torch.from_numpy(rng.integers(0, 360, size=(B, T)).astype(np.float32))
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    xy = torch.from_numpy(rng.uniform(-1, 1, size=(B, 2)).astype(np.float32))
    xy = xy / torch.clamp(xy.abs().max(dim=1, keepdim=True).values, min=1.0)
    lens = [int(x) for x in rng.integers(1, T + 1, size=B)]
    lens[0] = T                                            # max(lenths) must be T (enc_vl.py:44-48)
    return dict(corners_px=torch.from_numpy(corners.astype(np.int32)), lang=torch.randn(B, L, 768, generator=g),
                lang_cls=torch.relu(torch.randn(B, 49, generator=g)), directions=dirs.contiguous(),
                gt_xy=xy.contiguous(), gt_alt=torch.from_numpy(rng.uniform(0, 1, size=B).astype(np.float32)),
                gt_prog=torch.from_numpy(rng.uniform(0, 1, size=B).astype(np.float32)), lenths=lens)


def _to_dev(hb, perm=None):
    out = {}
    for k, v in hb.items():
        if torch.is_tensor(v):
            out[k] = (v if perm is None else v[perm]).contiguous().cuda()
        else:
            out[k] = list(v) if perm is None else [v[i] for i in perm.tolist()]
    return out


def test_config2_training_step_properties(built_lib):
    from avdn_b200.xview_et.agent import NavCMTAgent
    B, T, L = 64, 10, 250
    cfg = _cfg_file()
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=cfg, darknet_weight_file=None,
                                 lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2, no_dropout=True)
    torch.manual_seed(0)
    agent = NavCMTAgent(args, device="cuda:0")
    os.unlink(cfg)
    agent.renderer.add_map("tile", syn.synthetic_tile(seed=0, size=SIZE), syn.synthetic_attention_tile(seed=0, size=SIZE))
    hb = _train_batch(B, T, L, seed=3)
    # ---- forward + loss: per-sample losses add up to the step loss (agent.py:883-885) ----
    loss, output, h_sali = agent.forward_loss(_to_dev(hb))
    torch.cuda.synchronize()
    li = agent._ctx[2]["loss_i"].clone()
    out0 = output.clone()
    l0 = float(loss.item())                               # (loss is the agent's reused accumulator)
    assert torch.isfinite(li).all() and torch.isfinite(out0).all()
    assert abs(l0 - float(li.sum().item()) * 0.2 / B) <= 1e-9 * abs(l0) + 1e-12
    # ---- permuting the episodes permutes the per-sample results (eval BN: no cross-sample coupling) ----
    perm = torch.from_numpy(np.random.default_rng(1).permutation(B))
    loss_p, output_p, _ = agent.forward_loss(_to_dev(hb, perm))
    torch.cuda.synchronize()
    li_p = agent._ctx[2]["loss_i"].clone()
    assert torch.allclose(output_p, out0[perm.cuda()], rtol=1e-5, atol=1e-6)
    assert torch.allclose(li_p, li[perm.cuda()], rtol=1e-5, atol=1e-8)
    assert abs(float(loss_p.item()) - l0) <= 1e-5 * abs(l0)
    # ---- training steps at the full size: finite, and the loss goes down on a fixed batch ----
    for opt in agent.optimizers:
        opt.lr = 1e-4
    dev_b = _to_dev(hb)
    losses = [agent.train_step(dev_b, sync_loss=True) for _ in range(6)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses
    for opt in agent.optimizers:
        assert torch.isfinite(opt.p).all()


def test_config5_rollout_is_shard_invariant(built_lib):
    from avdn_b200.xview_lstm.agent import NavCMTAgent
    import bench_rollout as br
    B, T, L = 256, 20, 250
    cfg = _cfg_file()
    torch.manual_seed(0)
    agent = NavCMTAgent(types.SimpleNamespace(darknet_model_file=cfg, darknet_weight_file=None, max_action_len=T),
                        device="cuda:0")
    os.unlink(cfg)
    # non-trivial running statistics so that the eval-mode trunk is well conditioned and not degenerate
    g = torch.Generator(device="cuda").manual_seed(2)
    for m in agent.vision_model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_var.copy_(torch.rand(m.num_features, device="cuda", generator=g) * 0.5 + 0.75)
    agent.renderer.add_map("tile", syn.synthetic_tile(seed=0, size=SIZE), None)
    hb = br.synthetic_rollout_batch(B, L, seed=0)
    full = {k: v.clone() for k, v in agent.rollout_greedy({k: v.cuda() for k, v in hb.items()}, T).items()}
    torch.cuda.synchronize()
    assert full["corners"].shape == (T + 1, B, 4, 2) and torch.isfinite(full["corners"]).all()
    # stop flags are sticky and everybody has ended after the last step (agent.py:700-703)
    e = full["ended"].bool()
    assert bool(e[-1].all()) and bool((e[1:] | ~e[:-1]).all())
    for lo, hi in ((0, 64), (192, 256)):
        part = agent.rollout_greedy({k: v[lo:hi].contiguous().cuda() for k, v in hb.items()}, T)
        torch.cuda.synchronize()
        for key in ("angle", "altitude", "ended"):
            assert torch.equal(part[key], full[key][:, lo:hi]), (key, lo)
        assert torch.equal(part["output"], full["output"][:, lo:hi]), lo
        assert torch.equal(part["corners"], full["corners"][:, lo:hi]), lo
