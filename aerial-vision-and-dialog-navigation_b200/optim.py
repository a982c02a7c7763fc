"""Flat-arena optimiser of the training hot path.

The reference keeps three ``torch.optim.AdamW`` instances (ET, BERT, Darknet;
src/xview_et/agent.py:153-156) and clips the ET gradients to a global norm of 40
(``agent.py:247``).  Here the parameters of one model are re-homed into ONE
contiguous fp32 buffer (``p``) with matching ``g`` / ``m`` / ``v`` buffers:

* the backward kernels accumulate straight into ``g`` (no autograd ``.grad`` copies);
* gradient clipping is one ``avdn_sumsq`` launch, the update one ``avdn_adamw`` launch;
* data-parallel training all-reduces slices of ``g`` (one NCCL call per bucket).

Parameters without a gradient in the reference (``dec_action.*``,
``attention_layer_vision.c.*``: grad is None, so AdamW skips them and
``clip_grad_norm_`` ignores them) are simply not part of the arena.
"""
from __future__ import annotations

import torch

from . import _lib


class FusedAdamW:
    def __init__(self, named_params, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=None):
        named_params = list(named_params.items()) if isinstance(named_params, dict) else list(named_params)
        if not named_params:
            raise ValueError("no parameters")
        dev = named_params[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW lives on the CUDA device (no CPU fallback)")
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.offsets = {}
        off = 0
        for n, p in named_params:
            self.offsets[n] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4            # keep every view 16-byte aligned
        self.n = off
        f32 = torch.float32
        self.p = torch.zeros(off, dtype=f32, device=dev)
        self.g = torch.zeros(off, dtype=f32, device=dev)
        self.m = torch.zeros(off, dtype=f32, device=dev)
        self.v = torch.zeros(off, dtype=f32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.grads = {}
        self.params = {}
        for n, p in named_params:
            o, k = self.offsets[n]
            view = self.p[o:o + k].view(p.shape)
            view.copy_(p.data)
            p.data = view                                 # the module now reads the arena
            self.grads[n] = self.g[o:o + k].view(p.shape)
            self.params[n] = p
        self.step_count = 0

    def zero_grad(self):
        self.g.zero_()

    def expose_grads(self):
        """Make ``p.grad`` alias the arena (inspection / torch-style code)."""
        for n, p in self.params.items():
            p.grad = self.grads[n]

    def grad_norm(self):
        """Global L2 norm of the arena gradients (host float; synchronises)."""
        self.sumsq.zero_()
        _lib.call("avdn_sumsq", _lib.ptr(self.g), self.n, _lib.ptr(self.sumsq))
        return float(self.sumsq.sqrt().item())

    def step(self, grad_scale=1.0):
        """clip_grad_norm_(max_norm) (if set) + AdamW.  ``grad_scale`` multiplies the
        gradients first (1/world_size after a sum all-reduce).  Returns kernel launches."""
        self.step_count += 1
        ptr = _lib.ptr
        launches = 1
        ss = None
        if self.max_norm is not None:
            self.sumsq.zero_()
            _lib.call("avdn_sumsq", ptr(self.g), self.n, ptr(self.sumsq))
            ss = self.sumsq
            launches += 2
        _lib.call("avdn_adamw", ptr(self.p), ptr(self.g), ptr(self.m), ptr(self.v), self.n, self.lr, self.betas[0],
                  self.betas[1], self.eps, self.wd, self.step_count, ptr(ss),
                  float(self.max_norm if self.max_norm is not None else 0.0), float(grad_scale))
        return launches
