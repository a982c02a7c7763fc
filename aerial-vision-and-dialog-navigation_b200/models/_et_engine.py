"""Execution engine of the episodic transformer (ET) for one (B, L, T) shape.

Owns every activation / gradient buffer and every GEMM plan of the forward and
backward pass; the modules in ``ET_haa.py`` / ``enc_vl.py`` are parameter
containers with the reference's ``state_dict`` layout.  Data layout in HBM:

* residual stream            fp32 ``[B*S, 768]`` (S = L + 2T), bf16 shadow for GEMM operands
* q|k|v                      bf16 ``[B*S, 2304]`` (nn.MultiheadAttention's packed in_proj order)
* scores / dP                fp32 ``[B, H, S, Sp]`` (Sp = S rounded up to 64; shared by all layers)
* P (softmax), dS            bf16 ``[B, H, S, Sp]`` (P kept per layer for backward)
* weights                    bf16 shadows of the fp32 masters, refreshed every forward

Every dense contraction is one ``avdn_gemm`` plan (tcgen05); the masks are
predicates inside the softmax kernel.  Reference: src/models/ET_haa.py:121-184,
src/models/enc_vl.py:34-83, torch ``nn.TransformerEncoderLayer`` (post-norm, relu).
"""
from __future__ import annotations

import math

import torch

from .. import _lib
from .. import gemm as G

E = 768
NSP = 49
NCH = 512
LN_EPS = 1e-5


def _cdiv(a, b):
    return (a + b - 1) // b


class _LayerBufs:
    pass


class ETEngine:
    def __init__(self, params: dict, n_heads: int, n_layers: int, d_ff: int, B: int, L: int, T: int, device,
                 grads: dict | None = None, with_frame_attn=True, with_heads=True):
        """``params``: name -> fp32 tensor (reference state_dict names).  ``grads``:
        optional name -> fp32 tensor to accumulate parameter gradients into (the
        optimiser's flat arena); engine-owned buffers are used otherwise."""
        assert E % n_heads == 0 and E // n_heads == 64, "head dim must be 64 (768 / 12 heads)"
        assert d_ff % 64 == 0
        self.P = params
        self.H, self.NL, self.FF = n_heads, n_layers, d_ff
        self.B, self.L, self.T = B, L, T
        self.S = L + 2 * T
        self.Sp = _cdiv(self.S, 64) * 64
        self.M = B * self.S
        self.dev = device
        self.with_frame_attn, self.with_heads = with_frame_attn, with_heads
        self.launches = 0
        f32, bf = torch.float32, torch.bfloat16
        M, S, Sp, H = self.M, self.S, self.Sp, self.H

        def buf(shape, dt=f32, zero=False):
            return (torch.zeros if zero else torch.empty)(shape, dtype=dt, device=device)

        # ---- parameter gradients ----
        self.G = {}
        for n, p in params.items():
            if grads is not None and n in grads:
                self.G[n] = grads[n]
            else:
                self.G[n] = torch.zeros_like(p)
        # ---- frame attention ----
        if with_frame_attn:
            self.fa_attn = buf((B * T, NCH))
            self.fa_wc = buf((B * T, NSP))
            self.fa_e49 = buf((B * T, NSP))
            self.emb_frames = buf((B * T, E))
        self.lens = torch.zeros(B, dtype=torch.int32, device=device)
        self._lens_host = None
        # ---- encoder ----
        self.v0 = buf((M, E)); self.x0 = buf((M, E)); self.x0h = buf((M, E), bf)
        self.mean0 = buf(M); self.rstd0 = buf(M)
        self.scores = buf((B, H, S, Sp))
        self.layers = []
        for l in range(n_layers):
            Lb = _LayerBufs()
            Lb.qkv = buf((M, 3 * E), bf)
            Lb.Pm = buf((B, H, S, Sp), bf)
            Lb.ctx = buf((M, E), bf)
            Lb.v1 = buf((M, E)); Lb.x1 = buf((M, E)); Lb.x1h = buf((M, E), bf)
            Lb.mean1 = buf(M); Lb.rstd1 = buf(M)
            Lb.h = buf((M, d_ff), bf)
            Lb.v2 = buf((M, E)); Lb.x2 = buf((M, E)); Lb.x2h = buf((M, E), bf)
            Lb.mean2 = buf(M); Lb.rstd2 = buf(M)
            Lb.w_in = buf((3 * E, E), bf); Lb.w_o = buf((E, E), bf)
            Lb.w_1 = buf((d_ff, E), bf); Lb.w_2 = buf((E, d_ff), bf)
            self.layers.append(Lb)
        self.tmp_f32 = buf((M, E))             # attn-out / ffn-out before the residual LayerNorm
        # ---- heads ----
        if with_heads:
            self.h0 = buf((B, 256)); self.h1 = buf((B, 32))
            self.output = buf((B, 4)); self.h_sali = buf((B, 64))
        self._sig = None
        self._bwd_ready = False
        # train-mode dropout (nn.TransformerEncoderLayer p, heads p): off until set_dropout()
        self.p_enc, self.p_head, self.seed = 0.0, 0.0, 0
        self._drop_plans_p = None

    HEAD_SITE = 1000

    def set_dropout(self, p_enc=0.0, p_head=0.0, seed=0):
        """Dropout of the NEXT forward/backward pair: ``p_enc`` at the four sites of every encoder layer
        (attention probabilities, dropout1, FFN hidden, dropout2: sites 4l..4l+3), ``p_head`` in the heads
        (sites 1000..1002).  Masks are a stateless hash of (seed, site, element): nothing is stored."""
        self.p_enc, self.p_head, self.seed = float(p_enc), float(p_head), int(seed) & 0xFFFFFFFFFFFFFFFF
        if self.p_enc > 0.0:
            for Lb in self.layers:
                if getattr(Lb, "Pu", None) is None:
                    Lb.Pu = torch.empty_like(Lb.Pm)          # full softmax (the backward needs it)

    # ------------------------------------------------------------------ names
    @staticmethod
    def lp(l):
        return f"encoder_vl.enc_transformer.layers.{l}."

    def _signature(self):
        return tuple(self.P[n].data_ptr() for n in sorted(self.P))

    # ------------------------------------------------------------------ plans
    def _build_fwd_plans(self):
        B, H, S, Sp, M, FF = self.B, self.H, self.S, self.Sp, self.M, self.FF
        P = self.P
        for l, Lb in enumerate(self.layers):
            pre = self.lp(l)
            xin = self.x0h if l == 0 else self.layers[l - 1].x2h
            Lb.xin = xin
            Lb.p_qkv = G.plan_plain(M=M, N=3 * E, K=E, a_ptr=xin.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_in.data_ptr(),
                                    ldb=E, b_mn=0, out=Lb.qkv, ldc=3 * E, bias=P[pre + "self_attn.in_proj_bias"],
                                    keep=(xin, Lb.w_in))
            q, k, v = Lb.qkv.data_ptr(), Lb.qkv.data_ptr() + E * 2, Lb.qkv.data_ptr() + 2 * E * 2
            hs, bs = 64, S * 3 * E                                     # head / sample strides inside qkv
            Lb.p_scores = G.plan_plain(M=S, N=S, K=64, a_ptr=q, lda=3 * E, a_mn=0, b_ptr=k, ldb=3 * E, b_mn=0,
                                       out=self.scores, ldc=Sp, alpha=1.0 / math.sqrt(64.0), batch0=H, batch1=B,
                                       a_bs=(hs, bs), b_bs=(hs, bs), out_bs=(S * Sp, H * S * Sp), keep=(Lb.qkv,))
            Lb.p_pv = G.plan_plain(M=S, N=64, K=S, a_ptr=Lb.Pm.data_ptr(), lda=Sp, a_mn=0, b_ptr=v, ldb=3 * E,
                                   b_mn=1, out=Lb.ctx, ldc=E, batch0=H, batch1=B, a_bs=(S * Sp, H * S * Sp),
                                   b_bs=(hs, bs), out_bs=(64, S * E), keep=(Lb.Pm, Lb.qkv))
            Lb.p_o = G.plan_plain(M=M, N=E, K=E, a_ptr=Lb.ctx.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_o.data_ptr(),
                                  ldb=E, b_mn=0, out=self.tmp_f32, ldc=E, bias=P[pre + "self_attn.out_proj.bias"],
                                  keep=(Lb.ctx, Lb.w_o))
            Lb.p_ff1 = G.plan_plain(M=M, N=FF, K=E, a_ptr=Lb.x1h.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_1.data_ptr(),
                                    ldb=E, b_mn=0, out=Lb.h, ldc=FF, bias=P[pre + "linear1.bias"], relu=True,
                                    keep=(Lb.x1h, Lb.w_1))
            Lb.p_ff2 = G.plan_plain(M=M, N=E, K=FF, a_ptr=Lb.h.data_ptr(), lda=FF, a_mn=0, b_ptr=Lb.w_2.data_ptr(),
                                    ldb=FF, b_mn=0, out=self.tmp_f32, ldc=E, bias=P[pre + "linear2.bias"],
                                    keep=(Lb.h, Lb.w_2))

    def _build_bwd_plans(self):
        B, H, S, Sp, M, FF = self.B, self.H, self.S, self.Sp, self.M, self.FF
        f32, bf, dev = torch.float32, torch.bfloat16, self.dev
        Gd = self.G
        self.dx = torch.zeros((M, E), dtype=f32, device=dev)       # gradient entering the top of the stack
        self.dva = torch.empty((M, E), dtype=f32, device=dev)      # LN-backward outputs (fp32 residual branch)
        self.dvb = torch.empty((M, E), dtype=f32, device=dev)
        self.dvh = torch.empty((M, E), dtype=bf, device=dev)       # ... and their bf16 shadow (GEMM operand)
        self.dh = torch.empty((M, FF), dtype=bf, device=dev)
        self.dbranch = torch.empty((M, E), dtype=f32, device=dev)  # dgrad of the sub-block into the residual
        self.dctx = torch.empty((M, E), dtype=bf, device=dev)
        self.dqkv = torch.empty((M, 3 * E), dtype=bf, device=dev)
        self.dS = torch.empty((B, H, S, Sp), dtype=bf, device=dev)
        self.dv0 = torch.empty((M, E), dtype=f32, device=dev)
        if self.with_frame_attn:
            self.d_emb = torch.empty((B * self.T, E), dtype=f32, device=dev)
        sk = lambda gm, gn: max(1, min(_cdiv(M, 64), (2 * 148) // max(1, gm * gn)))
        for l, Lb in enumerate(self.layers):
            pre = self.lp(l)
            # ---- FFN ----
            Lb.b_ff2_d = G.plan_plain(M=M, N=FF, K=E, a_ptr=self.dvh.data_ptr(), lda=E, a_mn=0,
                                      b_ptr=Lb.w_2.data_ptr(), ldb=FF, b_mn=1, out=self.dh, ldc=FF, relu_mask=Lb.h,
                                      keep=(self.dvh, Lb.w_2))
            Lb.b_ff2_w = G.plan_plain(M=E, N=FF, K=M, a_ptr=self.dvh.data_ptr(), lda=E, a_mn=1,
                                      b_ptr=Lb.h.data_ptr(), ldb=FF, b_mn=1, out=Gd[pre + "linear2.weight"], ldc=FF,
                                      accumulate=2, split_k=sk(_cdiv(E, 128), _cdiv(FF, 256)),
                                      keep=(self.dvh, Lb.h))
            Lb.b_ff1_d = G.plan_plain(M=M, N=E, K=FF, a_ptr=self.dh.data_ptr(), lda=FF, a_mn=0,
                                      b_ptr=Lb.w_1.data_ptr(), ldb=E, b_mn=1, out=self.dbranch, ldc=E,
                                      keep=(self.dh, Lb.w_1))
            Lb.b_ff1_w = G.plan_plain(M=FF, N=E, K=M, a_ptr=self.dh.data_ptr(), lda=FF, a_mn=1,
                                      b_ptr=Lb.x1h.data_ptr(), ldb=E, b_mn=1, out=Gd[pre + "linear1.weight"], ldc=E,
                                      accumulate=2, split_k=sk(_cdiv(FF, 128), _cdiv(E, 256)),
                                      keep=(self.dh, Lb.x1h))
            # ---- attention output projection ----
            Lb.b_o_d = G.plan_plain(M=M, N=E, K=E, a_ptr=self.dvh.data_ptr(), lda=E, a_mn=0,
                                    b_ptr=Lb.w_o.data_ptr(), ldb=E, b_mn=1, out=self.dctx, ldc=E,
                                    keep=(self.dvh, Lb.w_o))
            Lb.b_o_w = G.plan_plain(M=E, N=E, K=M, a_ptr=self.dvh.data_ptr(), lda=E, a_mn=1,
                                    b_ptr=Lb.ctx.data_ptr(), ldb=E, b_mn=1, out=Gd[pre + "self_attn.out_proj.weight"],
                                    ldc=E, accumulate=2, split_k=sk(_cdiv(E, 128), _cdiv(E, 256)),
                                    keep=(self.dvh, Lb.ctx))
            # ---- attention core ----
            q, k, v = Lb.qkv.data_ptr(), Lb.qkv.data_ptr() + E * 2, Lb.qkv.data_ptr() + 2 * E * 2
            dq, dk, dv = self.dqkv.data_ptr(), self.dqkv.data_ptr() + E * 2, self.dqkv.data_ptr() + 2 * E * 2
            hs, bs = 64, S * 3 * E
            pbs = (S * Sp, H * S * Sp)
            # dP[q,k] = dctx[q,:] . V[k,:]
            Lb.b_dp = G.plan_plain(M=S, N=S, K=64, a_ptr=self.dctx.data_ptr(), lda=E, a_mn=0, b_ptr=v, ldb=3 * E,
                                   b_mn=0, out=self.scores, ldc=Sp, batch0=H, batch1=B, a_bs=(64, S * E),
                                   b_bs=(hs, bs), out_bs=pbs, keep=(self.dctx, Lb.qkv))
            # dV[k,:] = sum_q P[q,k] dctx[q,:]
            Lb.b_dv = G.plan_plain(M=S, N=64, K=S, a_ptr=Lb.Pm.data_ptr(), lda=Sp, a_mn=1,
                                   b_ptr=self.dctx.data_ptr(), ldb=E, b_mn=1, out=self.dqkv, out_ptr=dv,
                                   ldc=3 * E, batch0=H, batch1=B, a_bs=pbs, b_bs=(64, S * E), out_bs=(hs, bs),
                                   keep=(Lb.Pm, self.dctx))
            # dQ[q,:] = sum_k dS[q,k] K[k,:]
            Lb.b_dq = G.plan_plain(M=S, N=64, K=S, a_ptr=self.dS.data_ptr(), lda=Sp, a_mn=0, b_ptr=k, ldb=3 * E,
                                   b_mn=1, out=self.dqkv, out_ptr=dq, ldc=3 * E, batch0=H, batch1=B, a_bs=pbs,
                                   b_bs=(hs, bs), out_bs=(hs, bs), keep=(self.dS, Lb.qkv))
            # dK[k,:] = sum_q dS[q,k] Q[q,:]
            Lb.b_dk = G.plan_plain(M=S, N=64, K=S, a_ptr=self.dS.data_ptr(), lda=Sp, a_mn=1, b_ptr=q, ldb=3 * E,
                                   b_mn=1, out=self.dqkv, out_ptr=dk, ldc=3 * E, batch0=H, batch1=B, a_bs=pbs,
                                   b_bs=(hs, bs), out_bs=(hs, bs), keep=(self.dS, Lb.qkv))
            # ---- in_proj ----
            Lb.b_qkv_d = G.plan_plain(M=M, N=E, K=3 * E, a_ptr=self.dqkv.data_ptr(), lda=3 * E, a_mn=0,
                                      b_ptr=Lb.w_in.data_ptr(), ldb=E, b_mn=1, out=self.dbranch, ldc=E,
                                      keep=(self.dqkv, Lb.w_in))
            Lb.b_qkv_w = G.plan_plain(M=3 * E, N=E, K=M, a_ptr=self.dqkv.data_ptr(), lda=3 * E, a_mn=1,
                                      b_ptr=Lb.xin.data_ptr(), ldb=E, b_mn=1,
                                      out=Gd[pre + "self_attn.in_proj_weight"], ldc=E, accumulate=2,
                                      split_k=sk(_cdiv(3 * E, 128), _cdiv(E, 256)), keep=(self.dqkv, Lb.xin))
        self._bwd_ready = True

    def _ensure_drop_plans(self):
        """FFN-hidden dropout: d(relu+dropout) = relu_mask epilogue with alpha = 1/(1-p)."""
        if self._drop_plans_p == self.p_enc:
            return
        M, FF = self.M, self.FF
        for Lb in self.layers:
            Lb.b_ff2_d_drop = G.plan_plain(M=M, N=FF, K=E, a_ptr=self.dvh.data_ptr(), lda=E, a_mn=0,
                                           b_ptr=Lb.w_2.data_ptr(), ldb=FF, b_mn=1, out=self.dh, ldc=FF,
                                           relu_mask=Lb.h, alpha=1.0 / (1.0 - self.p_enc), keep=(self.dvh, Lb.w_2))
        self._drop_plans_p = self.p_enc

    def _ensure_plans(self):
        sig = self._signature()
        if sig != self._sig:
            self._build_fwd_plans()
            self._bwd_ready = False
            self._drop_plans_p = None
            self._sig = sig

    # ---------------------------------------------------------------- forward
    def _call(self, name, *a):
        _lib.call(name, *a)
        self.launches += 1

    def _run(self, plan):
        plan.run()
        self.launches += 1

    def set_lengths(self, lenths):
        lens = [int(x) for x in lenths]
        if len(lens) != self.B:
            raise ValueError(f"lenths has {len(lens)} entries for batch {self.B}")
        if max(lens) != self.T:
            raise ValueError(f"max(lenths)={max(lens)} must equal the number of frames T={self.T} "
                             "(the reference builds its masks from max(lenths), enc_vl.py:44-48)")
        if min(lens) < 1:
            raise ValueError("every episode needs at least one step")
        if lens != self._lens_host:
            self.lens.copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=False)
            self._lens_host = lens

    def refresh_weights(self):
        ptr = _lib.ptr
        for l, Lb in enumerate(self.layers):
            pre = self.lp(l)
            for name, dst in ((pre + "self_attn.in_proj_weight", Lb.w_in), (pre + "self_attn.out_proj.weight", Lb.w_o),
                              (pre + "linear1.weight", Lb.w_1), (pre + "linear2.weight", Lb.w_2)):
                self._call("avdn_cast_f32_bf16", ptr(self.P[name]), ptr(dst), dst.numel())

    def frame_attention(self, frames, lang_cls):
        """ET_haa.py:138-144 -> emb_frames [B*T,768]."""
        P, ptr = self.P, _lib.ptr
        self._call("avdn_frame_attn_fwd", ptr(frames), ptr(lang_cls), ptr(P["attention_layer_vision.linear_in.weight"]),
                   ptr(P["attention_layer_vision.linear_out.weight"]), ptr(P["fc2.weight"]), ptr(P["fc2.bias"]),
                   self.B, self.T, ptr(self.fa_attn), ptr(self.fa_wc), ptr(self.fa_e49), ptr(self.emb_frames))
        return self.emb_frames

    def encode(self, lang, emb_frames, dirs, embedded_dirs=False, pe=None):
        """enc_vl.py:34-83: returns the fp32 encoder output [B*S,768]."""
        P, ptr = self.P, _lib.ptr
        B, L, T, M, H = self.B, self.L, self.T, self.M, self.H
        self._ensure_plans()
        self.refresh_weights()
        wd = None if embedded_dirs else P["direction_embedding.weight"]
        bd = None if embedded_dirs else P["direction_embedding.bias"]
        self._call("avdn_embed_fwd", ptr(lang), ptr(emb_frames), ptr(dirs), ptr(wd), ptr(bd), ptr(pe), B, L, T,
                   ptr(self.v0))
        self._call("avdn_ln_fwd", ptr(self.v0), None, ptr(P["encoder_vl.enc_layernorm.weight"]),
                   ptr(P["encoder_vl.enc_layernorm.bias"]), M, E, LN_EPS, None, ptr(self.x0), ptr(self.x0h),
                   ptr(self.mean0), ptr(self.rstd0))
        x = self.x0
        for l, Lb in enumerate(self.layers):
            pre = self.lp(l)
            self._run(Lb.p_qkv)
            self._run(Lb.p_scores)
            pd, sd, st = self.p_enc, self.seed, 4 * l
            self._call("avdn_softmax_fwd_drop", ptr(self.scores), ptr(self.lens), B, H, L, T, self.Sp, ptr(Lb.Pm),
                       ptr(Lb.Pu) if pd > 0 else None, pd, sd, st)
            self._run(Lb.p_pv)
            self._run(Lb.p_o)
            self._call("avdn_ln_fwd_drop", ptr(x), ptr(self.tmp_f32), ptr(P[pre + "norm1.weight"]),
                       ptr(P[pre + "norm1.bias"]), M, E, LN_EPS, ptr(Lb.v1), ptr(Lb.x1), ptr(Lb.x1h), ptr(Lb.mean1),
                       ptr(Lb.rstd1), pd, sd, st + 1)
            self._run(Lb.p_ff1)
            if pd > 0:
                self._call("avdn_dropout_bf16", ptr(Lb.h), Lb.h.numel(), pd, sd, st + 2)
            self._run(Lb.p_ff2)
            self._call("avdn_ln_fwd_drop", ptr(Lb.x1), ptr(self.tmp_f32), ptr(P[pre + "norm2.weight"]),
                       ptr(P[pre + "norm2.bias"]), M, E, LN_EPS, ptr(Lb.v2), ptr(Lb.x2), ptr(Lb.x2h), ptr(Lb.mean2),
                       ptr(Lb.rstd2), pd, sd, st + 3)
            x = Lb.x2
        return x

    def heads(self, x):
        """ET_haa.py:157-166 -> output [B,4], h_sali [B,64]."""
        P, ptr = self.P, _lib.ptr
        d = "decoder_2_action_full."
        rv, rd = self.L + self.T - 1, self.L + 2 * self.T - 1
        self._call("avdn_heads_fwd_drop", ptr(x), self.B, self.S, rv, rd, ptr(P[d + "0.weight"]), ptr(P[d + "0.bias"]),
                   ptr(P[d + "3.weight"]), ptr(P[d + "3.bias"]), ptr(P[d + "6.weight"]), ptr(P[d + "6.bias"]),
                   ptr(P["fc.0.weight"]), ptr(P["fc.0.bias"]), ptr(self.h0), ptr(self.h1), ptr(self.output),
                   ptr(self.h_sali), self.p_head, self.seed, self.HEAD_SITE)
        return self.output, self.h_sali

    def forward(self, frames, lang, lang_cls, dirs, lenths, pe):
        """Full ET forward.  Tensors must stay alive/unmodified until backward()."""
        self.set_lengths(lenths)
        self._in = (frames, lang, lang_cls, dirs, pe)
        emb = self.frame_attention(frames, lang_cls)
        x = self.encode(lang, emb, dirs, False, pe)
        return self.heads(x)

    # --------------------------------------------------------------- backward
    def zero_grads(self):
        for g in self.G.values():
            g.zero_()

    def backward_heads(self, d_output, d_h_sali):
        P, Gd, ptr = self.P, self.G, _lib.ptr
        d = "decoder_2_action_full."
        x = self.layers[-1].x2
        rv, rd = self.L + self.T - 1, self.L + 2 * self.T - 1
        self.dx.zero_()
        self.launches += 1
        self._call("avdn_heads_bwd_drop", ptr(x), self.B, self.S, rv, rd, ptr(P[d + "0.weight"]),
                   ptr(P[d + "3.weight"]), ptr(P[d + "6.weight"]), ptr(P["fc.0.weight"]), ptr(self.h0), ptr(self.h1),
                   ptr(self.h_sali), ptr(d_output), ptr(d_h_sali), ptr(self.dx), ptr(Gd[d + "0.weight"]),
                   ptr(Gd[d + "0.bias"]), ptr(Gd[d + "3.weight"]), ptr(Gd[d + "3.bias"]), ptr(Gd[d + "6.weight"]),
                   ptr(Gd[d + "6.bias"]), ptr(Gd["fc.0.weight"]), ptr(Gd["fc.0.bias"]), self.p_head)

    def backward_encoder(self, d_out=None):
        """Backward through the transformer stack.  The gradient w.r.t. the encoder
        output is ``self.dx`` (filled by ``backward_heads``) or ``d_out`` [B*S,768].
        Returns dv0 = gradient w.r.t. the pre-LayerNorm embedding [B*S,768] fp32."""
        if not self._bwd_ready:
            self._build_bwd_plans_safe()
        P, Gd, ptr = self.P, self.G, _lib.ptr
        B, H, S, Sp, M, FF = self.B, self.H, self.S, self.Sp, self.M, self.FF
        BF, F32 = 0, 1
        if d_out is not None:
            self.dx.copy_(d_out.reshape(M, E))
            self.launches += 1
        pd, sd = self.p_enc, self.seed
        if pd > 0:
            self._ensure_drop_plans()
        dy1, dy2 = self.dx, None
        for l in reversed(range(self.NL)):
            Lb = self.layers[l]
            pre = self.lp(l)
            st = 4 * l
            # norm2 (dvh = gradient of the dropout2 branch)
            self._call("avdn_ln_bwd_drop", ptr(dy1), ptr(dy2), ptr(Lb.v2), ptr(Lb.mean2), ptr(Lb.rstd2),
                       ptr(P[pre + "norm2.weight"]), M, E, ptr(self.dva), ptr(self.dvh), ptr(Gd[pre + "norm2.weight"]),
                       ptr(Gd[pre + "norm2.bias"]), pd, sd, st + 3)
            # FFN
            self._call("avdn_colsum", ptr(self.dvh), BF, M, E, E, ptr(Gd[pre + "linear2.bias"]))
            self._run(Lb.b_ff2_w)
            self._run(Lb.b_ff2_d_drop if pd > 0 else Lb.b_ff2_d)
            self._call("avdn_colsum", ptr(self.dh), BF, M, FF, FF, ptr(Gd[pre + "linear1.bias"]))
            self._run(Lb.b_ff1_w)
            self._run(Lb.b_ff1_d)
            # norm1: dy = dva (residual) + dbranch (through the FFN)
            self._call("avdn_ln_bwd_drop", ptr(self.dva), ptr(self.dbranch), ptr(Lb.v1), ptr(Lb.mean1), ptr(Lb.rstd1),
                       ptr(P[pre + "norm1.weight"]), M, E, ptr(self.dvb), ptr(self.dvh), ptr(Gd[pre + "norm1.weight"]),
                       ptr(Gd[pre + "norm1.bias"]), pd, sd, st + 1)
            # attention
            self._call("avdn_colsum", ptr(self.dvh), BF, M, E, E, ptr(Gd[pre + "self_attn.out_proj.bias"]))
            self._run(Lb.b_o_w)
            self._run(Lb.b_o_d)
            self._run(Lb.b_dp)
            self._run(Lb.b_dv)
            self._call("avdn_softmax_bwd_drop", ptr(Lb.Pu if pd > 0 else Lb.Pm), ptr(self.scores), B * H * S, S, Sp,
                       1.0 / math.sqrt(64.0), ptr(self.dS), pd, sd, st)
            self._run(Lb.b_dq)
            self._run(Lb.b_dk)
            self._call("avdn_colsum", ptr(self.dqkv), BF, M, 3 * E, 3 * E, ptr(Gd[pre + "self_attn.in_proj_bias"]))
            self._run(Lb.b_qkv_w)
            self._run(Lb.b_qkv_d)
            dy1, dy2 = self.dvb, self.dbranch
        self._call("avdn_ln_bwd", ptr(dy1), ptr(dy2), ptr(self.v0), ptr(self.mean0), ptr(self.rstd0),
                   ptr(P["encoder_vl.enc_layernorm.weight"]), M, E, ptr(self.dv0), None,
                   ptr(Gd["encoder_vl.enc_layernorm.weight"]), ptr(Gd["encoder_vl.enc_layernorm.bias"]))
        return self.dv0

    def _build_bwd_plans_safe(self):
        self._ensure_plans()
        self._build_bwd_plans()

    def backward(self, d_output, d_h_sali, d_frames=None, need_lang_grad=False, d_lang_cls=None):
        """Full ET backward: accumulates parameter gradients into ``self.G`` and
        writes d_frames [B*T,512,49] (allocated if None).  Returns (d_frames, d_lang|None).
        ``d_lang_cls`` [B,49] fp32 (zero-filled by the caller) receives the gradient of ``lang_cls``
        (the BERT head's linear_cls, agent.py:527-543) when given."""
        P, Gd, ptr = self.P, self.G, _lib.ptr
        B, L, T, S = self.B, self.L, self.T, self.S
        frames, lang, lang_cls, dirs, pe = self._in
        if not self._bwd_ready:
            self._build_bwd_plans_safe()
        self.backward_heads(d_output, d_h_sali)
        dv0 = self.backward_encoder().view(B, S, E)
        self._call("avdn_embed_dir_bwd", ptr(dv0), ptr(dirs), B, L, T, ptr(Gd["direction_embedding.weight"]),
                   ptr(Gd["direction_embedding.bias"]))
        self.d_emb.view(B, T, E).copy_(dv0[:, L:L + T])
        self.launches += 1
        if d_frames is None:
            d_frames = torch.empty((B * T, NCH, NSP), dtype=torch.float32, device=self.dev)
        self._call("avdn_frame_attn_bwd_cls", ptr(frames), ptr(lang_cls),
                   ptr(P["attention_layer_vision.linear_in.weight"]),
                   ptr(P["attention_layer_vision.linear_out.weight"]), ptr(P["fc2.weight"]), B, T, ptr(self.fa_attn),
                   ptr(self.fa_wc), ptr(self.fa_e49), ptr(self.d_emb), ptr(d_frames),
                   ptr(Gd["attention_layer_vision.linear_in.weight"]),
                   ptr(Gd["attention_layer_vision.linear_out.weight"]), ptr(Gd["fc2.weight"]), ptr(Gd["fc2.bias"]),
                   ptr(d_lang_cls))
        self.launches += 1                       # fc2 weight-gradient reduction + the per-frame kernel
        d_lang = dv0[:, :L].contiguous() if need_lang_grad else None
        return d_frames, d_lang


class ETDecodeState:
    """Incremental ET inference for a greedy rollout (one per (B, L, Tmax)).

    The reference re-runs the whole encoder over the growing history at every step
    (src/xview_et/agent.py:620-627).  Rows of earlier steps cannot change: the attention mask is causal
    over steps (model_util.py:213-241: language rows see language only; frame/direction row t sees
    language and the frame/direction rows <= t) and LayerNorm / FFN act row by row.  So step 0 runs the
    full engine once (T = 1) and keeps every layer's K|V rows; step t >= 1 computes only its two new rows
    per sample (frame t, direction t) against the cache (``avdn_attn_decode``) -- the language rows, 93 %
    of the sequence, are encoded once per rollout instead of once per step.

    Cache layout per layer: bf16 ``[B, L + 2*Tmax, 1536]`` (k | v); row j < L = language token j, row
    L + 2s = frame s, row L + 2s + 1 = direction s (the T = 1 engine's row order for s = 0).
    Samples that have ended keep receiving rows (their outputs are ignored, as in the reference)."""

    R = 2

    def __init__(self, eng1: ETEngine, Tmax: int):
        assert eng1.T == 1 and eng1.with_frame_attn and eng1.with_heads
        self.e = eng1
        self.B, self.L, self.Tmax = eng1.B, eng1.L, Tmax
        self.Lc = self.L + 2 * Tmax
        B, dev, FF = self.B, eng1.dev, eng1.FF
        f32, bf = torch.float32, torch.bfloat16
        M = self.R * B
        self.M = M
        buf = lambda shape, dt=f32: torch.empty(shape, dtype=dt, device=dev)
        self.cache = [torch.zeros((B, self.Lc, 2 * E), dtype=bf, device=dev) for _ in eng1.layers]
        self.v0, self.x0, self.x0h = buf((M, E)), buf((M, E)), buf((M, E), bf)
        self.mean, self.rstd = buf(M), buf(M)
        self.qkv = buf((M, 3 * E), bf)
        self.ctx = buf((M, E), bf)
        self.tmp = buf((M, E))
        self.h = buf((M, FF), bf)
        self.xa = [(buf((M, E)), buf((M, E), bf)) for _ in eng1.layers]      # after norm1
        self.xb = [(buf((M, E)), buf((M, E), bf)) for _ in eng1.layers]      # after norm2
        self.launches = 0
        self._plans = None

    def _build_plans(self):
        e, M, FF, P = self.e, self.M, self.e.FF, self.e.P
        self._plans = []
        for l, Lb in enumerate(e.layers):
            pre = e.lp(l)
            xin = self.x0h if l == 0 else self.xb[l - 1][1]
            x1h, _ = self.xa[l][1], None
            pl = dict(
                qkv=G.plan_plain(M=M, N=3 * E, K=E, a_ptr=xin.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_in.data_ptr(), ldb=E,
                                 b_mn=0, out=self.qkv, ldc=3 * E, bias=P[pre + "self_attn.in_proj_bias"],
                                 keep=(xin, Lb.w_in)),
                o=G.plan_plain(M=M, N=E, K=E, a_ptr=self.ctx.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_o.data_ptr(), ldb=E,
                               b_mn=0, out=self.tmp, ldc=E, bias=P[pre + "self_attn.out_proj.bias"],
                               keep=(self.ctx, Lb.w_o)),
                ff1=G.plan_plain(M=M, N=FF, K=E, a_ptr=x1h.data_ptr(), lda=E, a_mn=0, b_ptr=Lb.w_1.data_ptr(), ldb=E,
                                 b_mn=0, out=self.h, ldc=FF, bias=P[pre + "linear1.bias"], relu=True,
                                 keep=(x1h, Lb.w_1)),
                ff2=G.plan_plain(M=M, N=E, K=FF, a_ptr=self.h.data_ptr(), lda=FF, a_mn=0, b_ptr=Lb.w_2.data_ptr(),
                                 ldb=FF, b_mn=0, out=self.tmp, ldc=E, bias=P[pre + "linear2.bias"],
                                 keep=(self.h, Lb.w_2)))
            self._plans.append(pl)
        self._sig = e._signature()

    def start(self):
        """After the T = 1 engine's forward of step 0: keep the K|V rows of every layer."""
        e = self.e
        S1 = self.L + 2
        for l, Lb in enumerate(e.layers):
            self.cache[l][:, :S1].copy_(Lb.qkv.view(self.B, S1, 3 * E)[:, :, E:])
            self.launches += 1

    def step(self, t, frames_t, dirs_t, lang_cls, pe):
        """Step t >= 1: ``frames_t`` [B,512,49] fp32, ``dirs_t`` [B,2] fp32 -> (output [B,4], h_sali [B,64])."""
        e, P, ptr, call = self.e, self.e.P, _lib.ptr, _lib.call
        B, L, M, H = self.B, self.L, self.M, e.H
        assert 1 <= t < self.Tmax
        if self._plans is None or self._sig != e._signature():
            self._build_plans()
        n0 = e.launches
        emb = e.frame_attention(frames_t, lang_cls)                      # [B,768] (T = 1 buffers)
        # the two new rows: frame t and direction t, both at position L + t (encodings.py:22-49)
        call("avdn_embed_fwd", None, ptr(emb), ptr(dirs_t), ptr(P["direction_embedding.weight"]),
             ptr(P["direction_embedding.bias"]), pe.data_ptr() + (L + t) * E * 4, B, 0, 1, ptr(self.v0))
        call("avdn_ln_fwd", ptr(self.v0), None, ptr(P["encoder_vl.enc_layernorm.weight"]),
             ptr(P["encoder_vl.enc_layernorm.bias"]), M, E, LN_EPS, None, ptr(self.x0), ptr(self.x0h), ptr(self.mean),
             ptr(self.rstd))
        x = self.x0
        row = L + 2 * t
        n = 2
        for l, pl in enumerate(self._plans):
            pre = e.lp(l)
            pl["qkv"].run()
            self.cache[l][:, row:row + 2].copy_(self.qkv.view(B, 2, 3 * E)[:, :, E:])
            call("avdn_attn_decode", ptr(self.qkv), ptr(self.cache[l]), B, self.R, H, self.Lc, row + 2,
                 1.0 / math.sqrt(64.0), ptr(self.ctx))
            pl["o"].run()
            x1, x1h = self.xa[l]
            call("avdn_ln_fwd", ptr(x), ptr(self.tmp), ptr(P[pre + "norm1.weight"]), ptr(P[pre + "norm1.bias"]), M, E,
                 LN_EPS, None, ptr(x1), ptr(x1h), ptr(self.mean), ptr(self.rstd))
            pl["ff1"].run()
            pl["ff2"].run()
            x2, x2h = self.xb[l]
            call("avdn_ln_fwd", ptr(x1), ptr(self.tmp), ptr(P[pre + "norm2.weight"]), ptr(P[pre + "norm2.bias"]), M, E,
                 LN_EPS, None, ptr(x2), ptr(x2h), ptr(self.mean), ptr(self.rstd))
            x = x2
            n += 9
        d = "decoder_2_action_full."
        call("avdn_heads_fwd", ptr(x), B, 2, 0, 1, ptr(P[d + "0.weight"]), ptr(P[d + "0.bias"]), ptr(P[d + "3.weight"]),
             ptr(P[d + "3.bias"]), ptr(P[d + "6.weight"]), ptr(P[d + "6.bias"]), ptr(P["fc.0.weight"]),
             ptr(P["fc.0.bias"]), ptr(e.h0), ptr(e.h1), ptr(e.output), ptr(e.h_sali))
        self.launches += n + 1 + (e.launches - n0)
        return e.output, e.h_sali
