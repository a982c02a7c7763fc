"""bench.py workload "bert": the language encoder of the AVDN loop -- CustomBERTModel forward +
backward at the ANDH shape (batch 64 dialogs x 250 tokens, bert-base geometry, random-init weights).  The metric
is dialogs/s; the roofline is the tcgen05 GEMM launches against the sustained bf16 peak."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

S_TOK = 250
VOCAB = 30522


def bert_flops_fwd(B, S, layers=12, d=768, ff=3072):
    # per token: qkv 3d^2 + out d^2 + ffn 2*d*ff MACs; attention 2*S*d MACs
    return 2 * B * S * layers * (4 * d * d + 2 * d * ff + 2 * S * d)


class BertWorkload:
    name = "bert_n1"
    metric = "CustomBERTModel train dialogs/s"
    unit = "dialogs/s"
    dtype = "bf16"
    B = 64
    CPU_SAMPLE = 2                    # dialogs per CPU pass; bench.py runs passes until N_CPU dialogs are done

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.B = int(os.environ.get("AVDN_BENCH_BATCH", self.B))

    def units_per_step(self):
        return self.B

    def config(self):
        return {"workload": "CustomBERTModel (bert-base: 12 layers x 768, 12 heads, ff 3072; + pooler + 2-layer head) "
                            "forward + backward, batch 64/GPU x 250 tokens with ragged right padding, random-init weights "
                            "(src/models/vln_model.py:128-159, call sites src/xview_et/agent.py:527-538)",
                "per_gpu_batch": self.B, "seq_len": S_TOK, "dropout": "train mode: 0.1 hidden / 0.1 attention / 0.2 head (stateless hash masks)",
                "cache": "activations of one pass (~7 GB) exceed L2; L2 is also flushed between steps",
                "parallelism": f"dialog-sharded x{self.world}, no collective (weights replicated, gradients local)"}

    def _inputs(self, B, seed):
        g = torch.Generator().manual_seed(seed)
        ids = torch.randint(0, VOCAB, (B, S_TOK), generator=g)
        lens = torch.randint(S_TOK // 3, S_TOK + 1, (B,), generator=g)
        lens[0] = S_TOK
        mask = (torch.arange(S_TOK)[None] < lens[:, None]).long()
        return ids, mask

    def setup_gpu(self, dev):
        from avdn_b200.models.bert import CustomBERTModel
        self.dev = dev
        torch.manual_seed(0)
        self.model = CustomBERTModel().to(dev).train()
        ids, mask = self._inputs(self.B, self.rank)
        self.pinned = (ids.pin_memory(), mask.pin_memory())
        self.ids, self.mask = ids.to(dev), mask.to(dev)
        self.eng = self.model.engine(self.B, S_TOK, dev)
        g = torch.Generator(device=dev).manual_seed(1)
        self.d_seq = torch.randn(self.B, S_TOK, 768, device=dev, generator=g) * 1e-3
        self.d_lin = torch.randn(self.B, 49, device=dev, generator=g) * 1e-3
        self.d_cls = torch.randn(self.B, 768, device=dev, generator=g) * 1e-3
        self.out_host = torch.empty((self.B, 49), dtype=torch.float32).pin_memory()
        self.profile = None

    def _fwd_bwd(self, ids, mask):
        e = self.eng
        e.set_dropout(*self.model.dropout_config())
        l0 = e.launches
        e.forward(ids, mask)
        e.zero_grads()
        e.backward(self.d_seq, self.d_lin, self.d_cls)
        return e.launches - l0

    def step(self):
        return self._fwd_bwd(self.ids, self.mask)

    def after_step(self, timed):
        pass

    def step_e2e(self):
        ids = self.pinned[0].to(self.dev, non_blocking=True)
        mask = self.pinned[1].to(self.dev, non_blocking=True)
        self._fwd_bwd(ids, mask)
        self.out_host.copy_(self.eng.lin, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(ids.numel() * 8 * 2), int(self.out_host.numel() * 4)

    def prepare_roofline(self):
        from avdn_b200 import _lib
        _lib.PROFILE = []
        self.step()
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
        return {"kernel": "gemm_kernel (tcgen05: projections, QK^T, PV, FFN and their backward)", "bound": "tensor",
                "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_sustained"] if ach else None, "traffic": None,
                "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)", "algorithmic_flops_per_step": g_fl,
                "analytic_flops_per_step": 3 * bert_flops_fwd(self.B, S_TOK), "gemm_ms_per_step": g_ms,
                "gemm_share_of_kernel_time": g_ms / tot if tot else None,
                "kernel_ms_breakdown": {k: {"n": v[0], "ms": round(v[1], 3)} for k, v in top}}

    # --------------------------------------------------------------------- CPU
    def cpu_step(self, n):
        """The third-party implementation the reference calls: transformers.BertModel (+ the head) on the host
        cores, forward + backward, CPU_SAMPLE dialogs x 250 tokens per pass."""
        if getattr(self, "_cpu", None) is None:
            from transformers import BertConfig, BertModel
            import torch.nn as nn
            torch.manual_seed(0)
            hf = BertModel(BertConfig())
            head = nn.Sequential(nn.Linear(768, 64), nn.ReLU(), nn.Dropout(0.2), nn.Linear(64, 49), nn.ReLU())
            hf.eval(); head.eval()
            self._cpu = (hf, head, self._inputs(self.CPU_SAMPLE, 0))
        hf, head, (ids, mask) = self._cpu
        done = 0
        while done < n:
            hf.zero_grad(); head.zero_grad()
            o = hf(ids, attention_mask=mask)
            lin = head(o["pooler_output"])
            (o["last_hidden_state"].square().mean() + lin.sum()).backward()
            done += self.CPU_SAMPLE
        return done

    def cpu_info(self):
        return {"kind": "reference", "cores": int(torch.get_num_threads()),
                "what": "transformers.BertModel + CustomBERTModel.linears (the library call of src/models/vln_model.py:"
                        "131,149), torch CPU fp32 forward + backward, 2 dialogs x 250 tokens per pass"}
