"""Rollout workload of bench.py: BASELINE.json configs[4] -- lstm_haa greedy waypoint-rollout inference,
batch 256, 20 steps, sharing the rendered-view + Darknet feature path.

A step = one full greedy rollout of 256 episodes: 20 x (GPS corners -> pixel corners -> homography -> 256
rendered views -> Darknet trunk in eval mode -> ViT_LSTM step -> discretise / stop test /
move_view_corners), all on the device with no host round trip inside the loop.  Sharded by episode
under torchrun (no collective)."""
from __future__ import annotations

import os
import tempfile
import types

import numpy as np
import torch

L_LANG, T_STEPS, SIZE = 250, 20, 3000
DARKNET_GFLOP_IMG = 15.295


def synthetic_rollout_batch(B, L, seed, size=SIZE):
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    span = 0.02
    lat_ratio = span / size
    bl, tr = np.array([40.0, -75.0]), np.array([40.0 + span, -75.0 + span])
    corners = np.zeros((B, 4, 2))
    dirs = np.zeros(B)
    for i in range(B):
        ctr = bl + span * rng.uniform(0.25, 0.75, size=2)
        half = span * rng.uniform(0.03, 0.09)                  # 180 .. 540 px views
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        corners[i] = ctr + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * half) @ R.T
        vec = (corners[i][0] + corners[i][1]) / 2 - corners[i].mean(0)
        ang = np.degrees(np.arctan2(vec[0], vec[1]))
        dirs[i] = round((360 - ang + 90) % 360) % 360
    geo = np.tile(np.array([bl[0], bl[1], tr[0], tr[1], lat_ratio]), (B, 1))
    return dict(corners_gps=torch.from_numpy(corners), directions=torch.from_numpy(dirs), geo=torch.from_numpy(geo),
                lang_feature=torch.randn(B, L, 768, generator=g), cls_hidden=torch.relu(torch.randn(B, 49, generator=g)))


class RolloutWorkload:
    name = "rollout_cfg5"
    metric = "lstm_haa greedy-rollout episodes/s"
    unit = "episodes/s"
    dtype = "bf16"
    B = 256
    CPU_SAMPLE = 4

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.B = int(os.environ.get("AVDN_BENCH_BATCH", self.B))

    def units_per_step(self):
        return self.B

    def config(self):
        return {"workload": "lstm_haa greedy waypoint-rollout inference, batch 256/GPU, 20 steps, 250-token dialog, views "
                            "rendered from a 3000x3000 tile, Darknet (eval BN) + ViT_LSTM + simulator update "
                            "(BASELINE configs[4])",
                "per_gpu_batch": self.B, "steps_per_rollout": T_STEPS, "views_per_rollout_per_gpu": self.B * T_STEPS,
                "cache": "activations of one trunk pass (6 GB at B=256) exceed L2; L2 is also flushed between rollouts",
                "parallelism": f"episode-sharded x{self.world}, no collective"}

    def setup_gpu(self, dev):
        from avdn_b200.utils import synthetic as mo
        from avdn_b200.utils import synthetic as wo
        from avdn_b200.xview_lstm.agent import NavCMTAgent
        self.dev = dev
        with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
            f.write(mo.yolov3_trunk_cfg())
        torch.manual_seed(0)
        self.agent = NavCMTAgent(types.SimpleNamespace(darknet_model_file=f.name, darknet_weight_file=None,
                                                       max_action_len=T_STEPS), device=dev)
        os.unlink(f.name)
        self.agent.renderer.add_map("tile", wo.synthetic_tile(seed=0, size=SIZE), None)
        self.host = synthetic_rollout_batch(self.B, L_LANG, seed=self.rank)
        self.pinned = {k: v.pin_memory() for k, v in self.host.items()}
        self.batch = {k: v.to(dev) for k, v in self.host.items()}
        self.batch["tile_idx"] = None
        self.res_host = torch.empty((T_STEPS + 1, self.B, 4, 2), dtype=torch.float64).pin_memory()
        self.profile = None

    def step(self):
        l0 = self.agent.launches
        self.agent.rollout_greedy(self.batch, T_STEPS)
        return self.agent.launches - l0

    def after_step(self, timed):
        pass

    def step_e2e(self):
        h2d = 0
        b = {"tile_idx": None}
        for k, v in self.pinned.items():
            b[k] = v.to(self.dev, non_blocking=True)
            h2d += v.numel() * v.element_size()
        res = self.agent.rollout_greedy(b, T_STEPS)
        self.res_host.copy_(res["corners"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h2d, int(self.res_host.numel() * 8)

    def prepare_roofline(self):
        from avdn_b200 import _lib
        _lib.PROFILE = []
        self.agent.rollout_greedy(self.batch, T_STEPS)
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl, nb in _lib.PROFILE:
            a = agg.setdefault(name, [0, 0.0, 0])
            a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
        _lib.PROFILE = None
        self.profile = agg

    def roofline(self, peaks):
        if self.profile is None:
            self.prepare_roofline()
        agg = self.profile
        g = {k: v for k, v in agg.items() if k.startswith("gemm")}
        g_ms, g_fl = sum(v[1] for v in g.values()), sum(v[2] for v in g.values())
        tot = sum(v[1] for v in agg.values())
        ach = g_fl / (g_ms * 1e-3) / 1e12 if g_ms else None
        top = sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]
        return {"kernel": "gemm_kernel (tcgen05 implicit-GEMM conv, forward only)", "bound": "tensor", "achieved": ach,
                "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"] if ach else None,
                "traffic": None, "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)",
                "algorithmic_flops_per_step": g_fl, "gemm_ms_per_step": g_ms,
                "gemm_share_of_kernel_time": g_ms / tot if tot else None,
                "kernel_ms_breakdown": {k: {"n": v[0], "ms": round(v[1], 3)} for k, v in top}}

    # --------------------------------------------------------------------- CPU
    def cpu_step(self, n):
        """Oracle port of the reference loop body (torch CPU fp32): trunk in eval mode + ViT_LSTM step + simulator
        update, CPU_SAMPLE episodes x 2 steps, scaled to a 20-step rollout."""
        from oracle import model_oracle as mo
        if getattr(self, "_cpu", None) is None:
            torch.manual_seed(0)
            cfg = mo.yolov3_trunk_cfg()
            sd = mo.random_trunk_state(cfg, seed=0)
            import torch.nn as nn
            lsd = {}
            for name, m in (("attention_layer_vision.linear_in", nn.Linear(49, 49, bias=False)),
                            ("attention_layer_vision.linear_out", nn.Linear(98, 49, bias=False)),
                            ("attention_layer_lang.linear_in", nn.Linear(768, 768, bias=False)),
                            ("attention_layer_lang.linear_out", nn.Linear(1536, 768, bias=False)),
                            ("vision_lstm", nn.LSTMCell(49, 576)), ("direct_lstm", nn.LSTMCell(32, 192)),
                            ("direction_embedding", nn.Linear(2, 32)), ("decoder_2_action_full.0", nn.Linear(768, 256)),
                            ("decoder_2_action_full.3", nn.Linear(256, 32)), ("decoder_2_action_full.6", nn.Linear(32, 4)),
                            ("fc.0", nn.Linear(49, 128)), ("fc.3", nn.Linear(128, 64))):
                for k, v in m.state_dict().items():
                    lsd[f"{name}.{k}"] = v.detach()
            hb = synthetic_rollout_batch(self.CPU_SAMPLE, L_LANG, seed=0)
            self._cpu = (mo, cfg, sd, lsd, hb, torch.randn(self.CPU_SAMPLE, 3, 224, 224))
        mo, cfg, sd, lsd, hb, images = self._cpu
        B = self.CPU_SAMPLE
        steps = 2
        done = 0.0
        with torch.no_grad():
            while done < n:
                state = None
                corners, dirs = hb["corners_gps"].numpy().copy(), hb["directions"].numpy().copy()
                ended = np.zeros(B, dtype=bool)
                for t in range(steps):
                    feats = mo.darknet_forward(images, sd, cfg, train=False).view(B, 512, 49)
                    r = mo.vit_lstm_step(lsd, feats, torch.from_numpy(dirs).long().view(B, 1), hb["cls_hidden"],
                                         hb["lang_feature"], state)
                    state = r[:4]
                    corners, dirs, ended, *_ = mo.waypoint_step(r[4].numpy(), corners, hb["geo"][:, :4].numpy(), dirs,
                                                                ended, 0.25, False)
                done += B * steps / T_STEPS
        return done

    def cpu_info(self):
        return {"kind": "port", "cores": int(torch.get_num_threads()),
                "what": "oracle/model_oracle.py (torch CPU fp32: Darknet eval forward + ViT_LSTM step + simulator update), "
                        "4 episodes x 2 of 20 steps per unit batch, scaled to whole rollouts (view rendering not included)"}


class ETRolloutWorkload(RolloutWorkload):
    """HAA-Transformer inference: greedy rollout of the ET agent (growing episode history), batch 64 per GPU,
    20 steps, never stopping early (stop threshold above 1) so that every rollout does the same work."""
    name = "et_rollout"
    metric = "et_haa greedy-rollout episodes/s"
    B = 64

    def config(self):
        return {"workload": "et_haa greedy waypoint-rollout inference (student feedback, src/xview_et/agent.py:580-760), "
                            "batch 64/GPU, 20 steps, 250-token dialog, views rendered from a 3000x3000 tile, Darknet "
                            "(eval BN) per step + ET over the growing history (1..20 steps) + simulator update; no early stop",
                "per_gpu_batch": self.B, "steps_per_rollout": T_STEPS, "views_per_rollout_per_gpu": self.B * T_STEPS,
                "cache": "L2 is flushed between rollouts",
                "parallelism": f"episode-sharded x{self.world}, no collective"}

    def setup_gpu(self, dev):
        from avdn_b200.utils import synthetic as syn
        from avdn_b200.xview_et.agent import NavCMTAgent
        from bench_train import make_args
        self.dev = dev
        with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
            f.write(syn.yolov3_trunk_cfg())
        torch.manual_seed(0)
        args = make_args(f.name)
        args.max_action_len = T_STEPS
        self.agent = NavCMTAgent(args, rank=self.rank, world_size=1, device=dev)
        os.unlink(f.name)
        self.agent.renderer.add_map("tile", syn.synthetic_tile(seed=0, size=SIZE), None)
        hb = synthetic_rollout_batch(self.B, L_LANG, seed=self.rank)
        self.host = dict(corners_gps=hb["corners_gps"], directions=hb["directions"], geo=hb["geo"],
                         lang=hb["lang_feature"], lang_cls=hb["cls_hidden"])
        self.pinned = {k: v.pin_memory() for k, v in self.host.items()}
        self.batch = {k: v.to(dev) for k, v in self.host.items()}
        self.batch["tile_idx"] = None
        self.res_host = torch.empty((T_STEPS + 1, self.B, 4, 2), dtype=torch.float64).pin_memory()
        self.profile = None

    def step(self):
        l0 = self.agent.launches
        self.agent.rollout_greedy(self.batch, T_STEPS, stop_threshold=2.0)
        return self.agent.launches - l0

    def step_e2e(self):
        h2d = 0
        b = {"tile_idx": None}
        for k, v in self.pinned.items():
            b[k] = v.to(self.dev, non_blocking=True)
            h2d += v.numel() * v.element_size()
        res = self.agent.rollout_greedy(b, T_STEPS, stop_threshold=2.0)
        self.res_host.copy_(res["corners"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h2d, int(self.res_host.numel() * 8)

    def prepare_roofline(self):
        from avdn_b200 import _lib
        _lib.PROFILE = []
        self.agent.rollout_greedy(self.batch, T_STEPS, stop_threshold=2.0)
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl, nb in _lib.PROFILE:
            a = agg.setdefault(name, [0, 0.0, 0])
            a[0] += 1; a[1] += e0.elapsed_time(e1); a[2] += fl
        _lib.PROFILE = None
        self.profile = agg

    def cpu_step(self, n):
        """Oracle port of the reference loop body (torch CPU fp32): trunk in eval mode + ET over the history +
        simulator update; CPU_SAMPLE episodes x the first 2 steps, scaled to a 20-step rollout by step count."""
        from oracle import model_oracle as mo
        if getattr(self, "_cpu", None) is None:
            cfg = mo.yolov3_trunk_cfg()
            hb = synthetic_rollout_batch(self.CPU_SAMPLE, L_LANG, seed=0)
            self._cpu = (cfg, mo.random_trunk_state(cfg, seed=0), mo.random_et_state(seed=0), hb,
                         torch.randn(self.CPU_SAMPLE, 3, 224, 224))
        cfg, sd, esd, hb, images = self._cpu
        B, steps, done = self.CPU_SAMPLE, 2, 0.0
        with torch.no_grad():
            while done < n:
                corners, dirs = hb["corners_gps"].numpy().copy(), hb["directions"].numpy().copy()
                ended = np.zeros(B, dtype=bool)
                fh, dh = [], []
                for t in range(steps):
                    fh.append(mo.darknet_forward(images, sd, cfg, train=False).view(B, 1, 512, 49))
                    rad = torch.from_numpy(dirs).float() / 180 * 3.14159
                    dh.append(torch.stack([torch.sin(rad), torch.cos(rad)], -1).view(B, 1, 2))
                    out, _, _ = mo.et_forward(esd, torch.cat(dh, 1), torch.cat(fh, 1), [t + 1] * B, hb["lang_feature"],
                                              hb["cls_hidden"])
                    corners, dirs, ended, *_ = mo.waypoint_step(out.numpy(), corners, hb["geo"][:, :4].numpy(), dirs,
                                                                ended, 2.0, False)
                done += B * steps / T_STEPS
        return done

    def cpu_info(self):
        return {"kind": "port", "cores": int(torch.get_num_threads()),
                "what": "oracle/model_oracle.py (torch CPU fp32: Darknet eval forward + ET over the history + simulator "
                        "update), 4 episodes x the first 2 of 20 steps per unit batch, scaled by step count (view "
                        "rendering not included; later steps attend over longer histories)"}
