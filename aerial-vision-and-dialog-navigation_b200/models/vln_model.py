"""``ViT_LSTM`` -- the recurrent policy of AVDN's LSTM baseline, mirror of
src/models/vln_model.py:163-250 (BASELINE config 5: greedy waypoint rollout).

Same constructor (``ViT_LSTM(args, vit_model)``), same sub-module names and therefore
the same ``state_dict`` keys (including the ``pos_embedding`` and
``attention_layer_vision_lang`` parameters that the reference's forward never uses),
same ``forward(current_direct, im_input, pos_input, cls_hidden, lang_feature, h_0, c_0,
hh_0, cc_0) -> (h_1, c_1, hh_1, cc_1, output, pred_saliency)``.

Inference only (the rollout of ``agent.test``, src/xview_lstm/agent.py:191-206, runs under
``.eval()``): dropout is the identity and the trunk uses its running statistics.  Training
the LSTM baseline is outside BASELINE's configs; ``.train()`` + forward raises.

Device work per step: trunk (tcgen05 convs, eval-mode BN folded into the apply pass),
``avdn_frame_attn_fwd`` (SoftDotAttention(49) over the 512 channels), ``avdn_linear_f32`` +
``avdn_lstm_cell`` for the two LSTM cells, ``avdn_lang_attn_fwd`` (SoftDotAttention(768) over
the dialog tokens) and fp32 linear layers for the two heads.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib

NCH, NSP, HID = 512, 49, 768


class SoftDotAttention(nn.Module):
    """Parameter container of src/models/vln_model.py:12-46."""

    def __init__(self, dim):
        super().__init__()
        self.linear_in = nn.Linear(dim, dim, bias=False)
        self.sm = nn.Softmax(dim=1)
        self.linear_out = nn.Linear(dim * 2, dim, bias=False)
        self.tanh = nn.Tanh()


class _StepBufs:
    """Per-(B, L) device buffers of one policy step.  The recurrent state lives here, double
    buffered by step parity: ``cat2[k]`` = [weighted (768) | h (192) | hh (576)] is at once the input of
    ``attention_layer_lang.linear_out`` and the home of the new hidden states."""

    def __init__(self, B, L, dev):
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.attn512, self.wc, self.e49 = f(B, NCH), f(B, NSP), f(B, NSP)
        self.gates_v, self.gates_d = f(B, 4 * 576), f(B, 4 * 192)
        self.deg, self.dir_emb = f(B), f(B, 32)
        self.cat2 = [f(B, 2 * HID), f(B, 2 * HID)]
        self.c_d, self.c_v = [f(B, 192), f(B, 192)], [f(B, 576), f(B, 576)]
        self.target, self.act_in = f(B, HID), f(B, HID)
        self.m0, self.m1, self.output = f(B, 256), f(B, 32), f(B, 4)
        self.s0, self.h_sali = f(B, 128), f(B, 64)
        self.k = 0                                   # parity of the NEXT step
        self.has_state = False

    def h(self, k):
        return self.cat2[k][:, HID:HID + 192]

    def hh(self, k):
        return self.cat2[k][:, HID + 192:]


class ViT_LSTM(nn.Module):
    def __init__(self, args, vit_model, hidden_size=768, dropout_ratio=0.5, im_channel_size=512, im_feature_size=49,
                 embedding_size=32):
        super().__init__()
        assert hidden_size == HID and im_channel_size == NCH and im_feature_size == NSP and embedding_size == 32, \
            "the reference hard-codes these sizes (vln_model.py:164-165,182,185)"
        self.args = args
        self.direction_embedding = nn.Linear(2, embedding_size)
        self.pos_embedding = nn.Linear(2, embedding_size)            # unused by forward (as in the reference)
        self.vision_model = vit_model
        self.attention_layer_lang = SoftDotAttention(hidden_size)
        self.attention_layer_vision_lang = SoftDotAttention(hidden_size)   # unused by forward
        self.attention_layer_vision = SoftDotAttention(im_feature_size)
        self.vision_lstm = nn.LSTMCell(im_feature_size, 576)
        self.drop = nn.Dropout(p=0.2)
        self.direct_lstm = nn.LSTMCell(embedding_size, 192)
        self.decoder_2_action_full = nn.Sequential(nn.Linear(hidden_size, 256), nn.ReLU(), nn.Dropout(0.2),
                                                   nn.Linear(256, 32), nn.ReLU(), nn.Dropout(0.2), nn.Linear(32, 4))
        self.fc = nn.Sequential(nn.Linear(im_feature_size, 128), nn.ReLU(), nn.Dropout(0.2), nn.Linear(128, 64),
                                nn.ReLU())
        self._bufs = {}
        self.launches = 0

    # ------------------------------------------------------------------ pieces
    def _linear(self, x, w, b, y, act=0, accumulate=0):
        M, K = x.shape
        N = w.shape[0]
        _lib.call("avdn_linear_f32", _lib.ptr(x), x.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(b), _lib.ptr(y),
                  y.stride(0), M, N, K, act, accumulate)
        self.launches += 1

    def _lstm(self, cell, x, h_prev, c_prev, gates, h_out, c_out):
        """nn.LSTMCell (vln_model.py:224-236): gates = W_ih x + b_ih + W_hh h + b_hh; zero state if None."""
        B, H = x.shape[0], cell.hidden_size
        if h_prev is None:                           # W_hh . 0 = 0: only the two bias vectors remain
            self._linear(x, cell.weight_ih, cell.bias_ih, gates)
            self._linear(torch.ones((B, 1), device=x.device), cell.bias_hh.view(-1, 1), None, gates, accumulate=1)
        else:
            self._linear(x, cell.weight_ih, cell.bias_ih, gates)
            self._linear(h_prev, cell.weight_hh, cell.bias_hh, gates, accumulate=1)
        _lib.call("avdn_lstm_cell", _lib.ptr(gates), _lib.ptr(c_prev), _lib.ptr(h_out), h_out.stride(0),
                  _lib.ptr(c_out), B, H)
        self.launches += 1

    def bufs(self, B, L, dev):
        key = (B, L, str(dev))
        bf = self._bufs.get(key)
        if bf is None:
            bf = self._bufs[key] = _StepBufs(B, L, dev)
        return bf

    def reset_state(self, B, L, dev):
        """Zero-init the recurrent state (src/xview_lstm/agent.py:546-550: h_0 = c_0 = hh_0 = cc_0 = None)."""
        bf = self.bufs(B, L, dev)
        bf.has_state, bf.k = False, 0
        return bf

    def step(self, im_feature, current_direct, cls_hidden, lang_feature, want_saliency=True):
        """One policy step on trunk features ``im_feature`` [B,512,49] f32 (device), advancing the
        recurrent state held in the (B, L) buffers.  Returns (output [B,4], h_sali [B,64] or None);
        both are views of step buffers, valid until the next call."""
        if self.training:
            raise NotImplementedError("ViT_LSTM is inference-only here (BASELINE config 5); call .eval()")
        _lib.require_cuda(im_feature, current_direct, cls_hidden, lang_feature)
        B, L = lang_feature.shape[0], lang_feature.shape[1]
        bf = self.bufs(B, L, im_feature.device)
        ptr, call = _lib.ptr, _lib.call
        k, p = bf.k, bf.k ^ 1
        prev = bf.has_state
        # named references: a temporary inside ptr(...) would be freed before the launch
        imf, cls, lf = im_feature.contiguous(), cls_hidden.contiguous(), lang_feature.contiguous()
        av, al = self.attention_layer_vision, self.attention_layer_lang
        # input_lstm_0 = SoftDotAttention(49)(cls_hidden, im_feature)              vln_model.py:219
        call("avdn_frame_attn_fwd", ptr(imf), ptr(cls), ptr(av.linear_in.weight),
             ptr(av.linear_out.weight), None, None, B, 1, ptr(bf.attn512), ptr(bf.wc), ptr(bf.e49), None)
        # hh_1, cc_1 = vision_lstm(drop(input_lstm_0), (hh_0, cc_0))                vln_model.py:220-226
        self._lstm(self.vision_lstm, bf.e49, bf.hh(p) if prev else None, bf.c_v[p] if prev else None, bf.gates_v,
                   bf.hh(k), bf.c_v[k])
        # direction_embedding([sin, cos](current_direct / 180 * 3.14159))           vln_model.py:228-229
        bf.deg.copy_(current_direct.reshape(B))      # int64 / float -> float32, as torch's true division does
        call("avdn_direction_embed", ptr(bf.deg), ptr(self.direction_embedding.weight),
             ptr(self.direction_embedding.bias), ptr(bf.dir_emb), B, 32)
        self._lstm(self.direct_lstm, bf.dir_emb, bf.h(p) if prev else None, bf.c_d[p] if prev else None, bf.gates_d,
                   bf.h(k), bf.c_d[k])
        # action_module_input = SoftDotAttention(768)(cat(h_1, hh_1), lang_feature)  vln_model.py:238-239
        hcat = bf.cat2[k][:, HID:]
        self._linear(hcat, al.linear_in.weight, None, bf.target)
        call("avdn_lang_attn_fwd", ptr(lf), ptr(bf.target), B, L, HID, None, ptr(bf.cat2[k]),
             bf.cat2[k].stride(0))
        self._linear(bf.cat2[k], al.linear_out.weight, None, bf.act_in, act=2)
        # output = decoder_2_action_full(action_module_input)                       vln_model.py:248
        dec = self.decoder_2_action_full
        self._linear(bf.act_in, dec[0].weight, dec[0].bias, bf.m0, act=1)
        self._linear(bf.m0, dec[3].weight, dec[3].bias, bf.m1, act=1)
        self._linear(bf.m1, dec[6].weight, dec[6].bias, bf.output)
        self.launches += 3
        h_sali = None
        if want_saliency:                                                          # vln_model.py:244
            self._linear(bf.e49, self.fc[0].weight, self.fc[0].bias, bf.s0, act=1)
            self._linear(bf.s0, self.fc[3].weight, self.fc[3].bias, bf.h_sali, act=1)
            h_sali = bf.h_sali
        bf.k, bf.has_state = p, True
        return bf.output, h_sali

    def state(self, B, L, dev):
        """(h, c, hh, cc) after the last step (copies)."""
        bf = self.bufs(B, L, dev)
        k = bf.k ^ 1
        return bf.h(k).clone(), bf.c_d[k].clone(), bf.hh(k).clone(), bf.c_v[k].clone()

    def forward(self, current_direct, im_input, pos_input, cls_hidden, lang_feature, h_0=None, c_0=None, hh_0=None,
                cc_0=None):
        """Reference signature (vln_model.py:213).  ``im_input`` [B,3,224,224] f32 normalised."""
        with torch.no_grad():
            feat = self.vision_model(im_input)                                      # [B,512,7,7]
            B, L = feat.shape[0], lang_feature.shape[1]
            given = [x is not None for x in (h_0, c_0, hh_0, cc_0)]
            if any(given) and not all(given):
                raise ValueError("pass all four recurrent states or none")
            bf = self.reset_state(B, L, feat.device)
            if all(given):                         # load the caller's state into the 'previous' buffers
                p = bf.k ^ 1
                bf.h(p).copy_(h_0); bf.c_d[p].copy_(c_0); bf.hh(p).copy_(hh_0); bf.c_v[p].copy_(cc_0)
                bf.has_state = True
            output, h_sali = self.step(feat.view(B, NCH, NSP), current_direct, cls_hidden, lang_feature)
            pred = torch.empty((B, 1, 224, 224), dtype=torch.float32, device=feat.device)
            _lib.call("avdn_upsample_saliency", _lib.ptr(h_sali), B, _lib.ptr(pred))
            h1, c1, hh1, cc1 = self.state(B, L, feat.device)
        return h1, c1, hh1, cc1, output.clone(), pred
