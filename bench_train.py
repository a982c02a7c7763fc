"""Training workload of bench.py: BASELINE.json configs[1] (and configs[3] under
torchrun): et_haa training step, bf16 tensor-core path, batch 64 per GPU,
synthetic ANDH shape (250-token dialog history, 10 views of 224x224 per
episode), Darknet features + transformer + both heads, loss, backward, AdamW.

A step = render the 640 views of the batch from the synthetic 3000x3000 tile,
Darknet trunk forward (train-mode BatchNorm), ET forward, fused loss, ET
backward, Darknet backward, (gradient all-reduce), clip + AdamW on both models.
"""
from __future__ import annotations

import math
import json
import os
import time
import types

import numpy as np
import torch

L_LANG, T_STEPS, SIZE = 250, 10, 3000
DARKNET_GFLOP_IMG = 15.295          # forward, per 224x224 image


def et_flops_fwd(S=270, d=768, layers=2):
    # per layer: qkv 3d^2 + out d^2 + ffn 2d^2 = 6 d^2 MACs per token; attention 2 S d MACs per token
    return layers * (2 * S * d * 6 * d + 4 * S * S * d)


def make_args(cfg_path):
    return types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=cfg_path,
                                 darknet_weight_file=None, lr=1e-5, nss_w=0.1, nss_r=0, ml_weight=0.2)


def synthetic_batch(B, T, L, seed, tile_size=SIZE):
    """Host (numpy / torch CPU) episode tensors of the ANDH shape."""
    from avdn_b200.utils import synthetic as wo   # synthetic inputs (product-side generators)
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    corners = wo.synthetic_pose_corners(B * T, seed=seed, size=tile_size, edge_frac=0.05).reshape(B, T, 4, 2)
    deg = torch.from_numpy(rng.integers(0, 360, size=(B, T)).astype(np.float32))
    dirs = torch.stack([torch.sin(deg / 180 * 3.14159), torch.cos(deg / 180 * 3.14159)], -1)
    xy = torch.from_numpy(rng.uniform(-1, 1, size=(B, 2)).astype(np.float32))
    xy = xy / torch.clamp(xy.abs().max(dim=1, keepdim=True).values, min=1.0)
    return dict(
        corners_px=torch.from_numpy(corners.astype(np.int32)),
        lang=torch.randn(B, L, 768, generator=g),
        lang_cls=torch.relu(torch.randn(B, 49, generator=g)),
        directions=dirs.contiguous(),
        gt_xy=xy.contiguous(),
        gt_alt=torch.from_numpy(rng.uniform(0, 1, size=B).astype(np.float32)),
        gt_prog=torch.from_numpy(rng.uniform(0, 1, size=B).astype(np.float32)),
        lenths=[T] * B,
    )


class TrainWorkload:
    name = "train_cfg2"
    metric = "HAA-Transformer train episodes/s"
    unit = "episodes/s"
    dtype = "bf16"
    B = 64
    CPU_SAMPLE = 4                    # episodes per CPU step = BASELINE configs[0] (batch 4)

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.B = int(os.environ.get("AVDN_BENCH_BATCH", self.B))

    def units_per_step(self):
        return self.B

    def config(self):
        return {"workload": "et_haa training step bf16, batch 64/GPU, synthetic ANDH shape: 250-token dialog, "
                            "10 views 224x224/episode rendered from a 3000x3000 tile, Darknet(57 conv, train-mode BN) "
                            "+ ET(2x768, 12 heads) + waypoint & human-attention heads, loss, backward, clip, AdamW "
                            "(BASELINE configs[1]; configs[3] under torchrun)",
                "global_batch": self.B * self.world, "per_gpu_batch": self.B, "views_per_step_per_gpu": self.B * T_STEPS,
                "seq_len": L_LANG + 2 * T_STEPS, "dropout": "train mode: 0.1 at the 4 sites of each encoder layer, 0.2 in the heads (stateless hash masks)",
                "cache": "per-step working set (~40 GB of activations) far exceeds L2; L2 is also flushed between steps",
                "parallelism": f"dp{self.world} by episode, NCCL all-reduce of gradients" if self.world > 1 else "dp1"}

    # --------------------------------------------------------------------- GPU
    def setup_gpu(self, dev):
        import tempfile
        from avdn_b200.utils import synthetic as mo
        from avdn_b200.utils import synthetic as wo
        from avdn_b200.xview_et.agent import NavCMTAgent
        self.dev = dev
        with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
            f.write(mo.yolov3_trunk_cfg())
        torch.manual_seed(0)
        self.agent = NavCMTAgent(make_args(f.name), rank=self.rank, world_size=self.world, device=dev)
        os.unlink(f.name)
        tile = wo.synthetic_tile(seed=0, size=SIZE)
        att = wo.synthetic_attention_tile(seed=0, size=SIZE)
        self.agent.renderer.add_map("tile", tile, att)
        self.host = synthetic_batch(self.B, T_STEPS, L_LANG, seed=self.rank)
        self.pinned = {k: v.pin_memory() for k, v in self.host.items() if torch.is_tensor(v)}
        self.batch = {k: v.to(dev) for k, v in self.host.items() if torch.is_tensor(v)}
        self.batch["lenths"] = self.host["lenths"]
        self.loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()
        self.profile = None
        self.last_loss = None

    def step(self):
        l0 = self.agent.launches
        self.agent.train_step(self.batch)
        return self.agent.launches - l0

    def after_step(self, timed):
        pass

    def step_e2e(self):
        h2d = 0
        b = {}
        for k, v in self.pinned.items():
            b[k] = v.to(self.dev, non_blocking=True)
            h2d += v.numel() * v.element_size()
        b["lenths"] = self.host["lenths"]
        loss = self.agent.train_step(b)
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.last_loss = float(self.loss_host.item())
        return h2d, 8

    def profile_step(self):
        """One instrumented step (outside the timed region): per-kernel CUDA-event times."""
        from avdn_b200 import _lib
        _lib.PROFILE = []
        self.agent.train_step(self.batch)
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl, nb in _lib.PROFILE:
            a = agg.setdefault(name, [0, 0.0, 0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += fl
        _lib.PROFILE = None
        self.profile = agg
        return agg

    def prepare_roofline(self):
        """Called on EVERY rank: the instrumented step is a full data-parallel step (all-reduce inside)."""
        if self.profile is None:
            self.profile_step()

    def flops_per_step(self):
        return (DARKNET_GFLOP_IMG * 1e9 * self.B * T_STEPS + et_flops_fwd() * self.B) * 3

    def roofline(self, peaks, ms_per_step=None):
        if self.profile is None:
            self.profile_step()
        g = {k: v for k, v in self.profile.items() if k.startswith("gemm")}
        g_ms = sum(v[1] for v in g.values())
        g_fl = sum(v[2] for v in g.values())
        g_n = sum(v[0] for v in g.values())
        tot = sum(v[1] for v in self.profile.values())
        ach = g_fl / (g_ms * 1e-3) / 1e12 if g_ms else None
        peak = peaks["bf16_sustained"]
        self._peak = peak
        top = sorted(self.profile.items(), key=lambda kv: -kv[1][1])[:14]
        # DRAM bytes of the step's gemm_kernel launches from the committed ncu capture of this same command
        # (profiles/r01_train_step_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum, single GPU)
        traffic, traffic_src = None, None
        tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_train_step_traffic.json")
        if self.world == 1 and os.path.exists(tp):
            try:
                with open(tp) as f:
                    tj = json.load(f)["gemm_kernel"]
                if tj["n"] == g_n:
                    traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
                    traffic_src = "profiles/r01_train_step_traffic.json (ncu, sum over the step's gemm_kernel launches)"
            except (KeyError, ValueError):
                pass
        return {"kernel": "gemm_kernel (tcgen05 implicit-GEMM conv fwd/dgrad/wgrad + transformer GEMMs)",
                "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peaks["source"] + " (sustained bf16 cuBLAS; kernel timed inside a long step)",
                "algorithmic_flops_per_step": g_fl, "gemm_launches_per_step": g_n, "gemm_ms_per_step": g_ms,
                "gemm_share_of_kernel_time": g_ms / tot if tot else None,
                "kernel_ms_breakdown": {k: {"n": v[0], "ms": round(v[1], 3)} for k, v in top}}

    def extra(self, ms_per_step=None):
        out = {"last_loss": self.last_loss}
        if ms_per_step:
            fl = self.flops_per_step()
            out["step_tflops_algorithmic"] = fl / 1e12
            out["step_tflops_per_s"] = fl / (ms_per_step * 1e-3) / 1e12
            out["mfu_vs_sustained_bf16"] = out["step_tflops_per_s"] / self._peak if getattr(self, "_peak", None) else None
        return out

    # --------------------------------------------------------------------- CPU
    def _cpu_setup(self):
        if getattr(self, "_cpu", None) is not None:
            return self._cpu
        from oracle import model_oracle as mo
        torch.manual_seed(0)
        B, T, L = self.CPU_SAMPLE, T_STEPS, L_LANG
        cfg = mo.yolov3_trunk_cfg()
        sd = mo.random_trunk_state(cfg, seed=0)
        et_sd = mo.random_et_state(seed=0)
        for d in (sd, et_sd):
            for k, v in d.items():
                if v.is_floating_point() and "running" not in k and not k.endswith(".pe"):
                    v.requires_grad_(True)
        hb = synthetic_batch(B, T, L, seed=0)
        images = torch.randn(B * T, 3, 224, 224)
        gt_sal = torch.zeros(B, 224, 224, dtype=torch.float64)
        gt_sal[:, 60:120, 80:160] = 1.0
        self._cpu = (mo, cfg, sd, et_sd, hb, images, gt_sal)
        return self._cpu

    def cpu_step(self, n):
        """The oracle's restatement of the reference step (torch CPU fp32 autograd: Darknet
        train-mode forward, ET forward, loss, backward) on ``CPU_SAMPLE`` episodes."""
        mo, cfg, sd, et_sd, hb, images, gt_sal = self._cpu_setup()
        B, T = self.CPU_SAMPLE, T_STEPS
        done = 0
        while done < n:
            for d in (sd, et_sd):
                for v in d.values():
                    v.grad = None
            feats = mo.darknet_forward(images, sd, cfg, train=True).view(B, T, 512, 49)
            out, sal, _ = mo.et_forward(et_sd, hb["directions"], feats, hb["lenths"], hb["lang"], hb["lang_cls"])
            loss = mo.step_loss(mo.et_loss(out, sal, hb["gt_xy"], hb["gt_alt"], hb["gt_prog"], gt_sal, 0.1), 0.2, B)
            loss.backward()
            done += B
        return done

    def cpu_info(self):
        return {"kind": "port", "cores": int(torch.get_num_threads()),
                "what": "oracle/model_oracle.py (torch CPU fp32 restatement of Darknet + ET + loss, fwd+bwd), "
                        "BASELINE configs[0] shape (batch 4, 10 views, 250 tokens)"}


class TrainRolloutWorkload(TrainWorkload):
    """The teacher-forced training rollout (src/xview_et/agent.py:580-760): the loss at every one of the 10 steps on
    the growing history -- views and trunk once over the 640 poses, then the reference's 10 encoder calls (history
    1..10), each with its loss and backward -- one optimiser step."""
    name = "train_rollout"
    metric = "HAA-Transformer train rollout episodes/s"

    def config(self):
        c = super().config()
        c["workload"] = c["workload"].replace("et_haa training step bf16", "et_haa teacher-forced training rollout "
                                              "(loss at each of the 10 steps on the history so far; 10 encoder "
                                              "forward+backward passes per step), bf16")
        return c

    def setup_gpu(self, dev):
        super().setup_gpu(dev)
        rng = np.random.default_rng(200 + self.rank)
        B, T = self.B, T_STEPS
        xy = torch.from_numpy(rng.uniform(-1, 1, size=(B, T, 2)).astype(np.float32))
        self.host.update(gt_xy=(xy / torch.clamp(xy.abs().amax(dim=2, keepdim=True), min=1.0)).contiguous(),
                         gt_alt=torch.from_numpy(rng.uniform(0, 1, size=(B, T)).astype(np.float32)),
                         gt_prog=torch.from_numpy(rng.uniform(0, 1, size=(B, T)).astype(np.float32)))
        self.host.pop("lenths")
        self.pinned = {k: v.pin_memory() for k, v in self.host.items() if torch.is_tensor(v)}
        self.batch = {k: v.to(dev) for k, v in self.host.items() if torch.is_tensor(v)}

    def step(self):
        l0 = self.agent.launches
        self.agent.train_rollout_step(self.batch)
        return self.agent.launches - l0

    def step_e2e(self):
        h2d = 0
        b = {}
        for k, v in self.pinned.items():
            b[k] = v.to(self.dev, non_blocking=True)
            h2d += v.numel() * v.element_size()
        loss = self.agent.train_rollout_step(b)
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.last_loss = float(self.loss_host.item())
        return h2d, 8

    def profile_step(self):
        from avdn_b200 import _lib
        _lib.PROFILE = []
        self.agent.train_rollout_step(self.batch)
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, fl, nb in _lib.PROFILE:
            a = agg.setdefault(name, [0, 0.0, 0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += fl
        _lib.PROFILE = None
        self.profile = agg
        return agg

    def flops_per_step(self):
        et = sum(et_flops_fwd(S=L_LANG + 2 * t) for t in range(1, T_STEPS + 1))
        return (DARKNET_GFLOP_IMG * 1e9 * self.B * T_STEPS + et * self.B) * 3

    def roofline(self, peaks, ms_per_step=None):
        r = super().roofline(peaks, ms_per_step)
        r["traffic"], r["traffic_source"] = None, None
        return r

    def cpu_step(self, n):
        raise SystemExit("the train_rollout workload has no CPU leg: use --no-cpu-baseline")


class TrainBertWorkload(TrainWorkload):
    """The training step with the language encoder in the loop (src/xview_et/agent.py:125-126,155,249,527-543):
    token ids in, BERT-base forward + backward + AdamW inside the step."""
    name = "train_bert"
    metric = "HAA-Transformer + BERT train episodes/s"

    def config(self):
        c = super().config()
        c["workload"] = c["workload"].replace("et_haa training step bf16", "et_haa training step with the language "
                                              "encoder (CustomBERTModel, bert-base, 250 tokens) trained in the loop, bf16")
        return c

    def setup_gpu(self, dev):
        super().setup_gpu(dev)
        self.agent.attach_lang_model()
        g = torch.Generator().manual_seed(100 + self.rank)
        ids = torch.randint(0, 30522, (self.B, L_LANG), generator=g)
        lens = torch.randint(L_LANG // 3, L_LANG + 1, (self.B,), generator=g)
        lens[0] = L_LANG
        mask = (torch.arange(L_LANG)[None] < lens[:, None]).long()
        for d in (self.host, ):
            d.pop("lang"); d.pop("lang_cls")
            d["input_ids"], d["attention_mask"] = ids, mask
        self.pinned = {k: v.pin_memory() for k, v in self.host.items() if torch.is_tensor(v)}
        self.batch = {k: v.to(dev) for k, v in self.host.items() if torch.is_tensor(v)}
        self.batch["lenths"] = self.host["lenths"]

    def flops_per_step(self):
        from bench_bert import bert_flops_fwd
        return super().flops_per_step() + 3 * bert_flops_fwd(self.B, L_LANG)

    def roofline(self, peaks, ms_per_step=None):
        r = super().roofline(peaks, ms_per_step)
        r["traffic"], r["traffic_source"] = None, None
        return r

    def cpu_step(self, n):
        raise SystemExit("the train_bert workload has no CPU leg: use --no-cpu-baseline (train and bert have one each)")
