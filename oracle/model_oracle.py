"""CPU (torch fp32 / fp64) restatement of stages 2 and 3 of the AVDN hot path.

TEST INFRASTRUCTURE ONLY — never imported by the product package.

Every function takes the reference's ``state_dict`` tensors by their reference
names, so the same random-init weights drive the reference modules, this oracle
and the CUDA path.  Reference ``file:line`` (relative to /root/reference):

* ``darknet_forward``            src/models/dark_net.py:17-33, 212-240
* ``soft_dot_attention``         src/models/ET_haa.py:54-74
* ``pos_encoding_table``         src/models/encodings.py:12-20
* ``mask_pad`` / ``attention_mask``  src/models/enc_vl.py:44-55, src/models/model_util.py:204-241
* ``encoder_vl_forward``         src/models/enc_vl.py:34-83 + torch ``nn.TransformerEncoderLayer``
                                 (post-norm, relu, eps 1e-5; third-party: torch, README.md:78)
* ``et_forward``                 src/models/ET_haa.py:121-184
* ``et_loss``                    src/xview_et/agent.py:663-681, 256-270, 883-885
* ``postprocess_waypoints``      src/xview_et/agent.py:637-653, 745-752
* ``vit_lstm_step``              src/models/vln_model.py:213-250 (config 5)
* ``move_view_corners`` / ``get_direction`` / ``waypoint_step``
                                 src/xview_et/agent.py:83-101, 285-384; src/xview_lstm/agent.py:607-626, 700-730

Parity pin: the reference has no tests or golden vectors.  This
restatement is pinned by executing the reference's own modules in the build
container on shared seeds/state_dicts (``tests/golden/make_model_golden.py``,
which also wrote the committed fixtures ``tests/golden/model_golden.pt``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PI_REF = 3.14159          # the reference's pi (agent.py:606,666,745)


# --------------------------------------------------------------------------
# Darknet
# --------------------------------------------------------------------------
def yolov3_trunk_cfg(truncate_after=None):
    """Text of the (inferred) truncated xview-yolov3 cfg: Darknet-53 + 5 head
    convs, ending at the 512-channel stride-32 layer."""
    out = ["[net]", "channels=3", "height=416", ""]

    def conv(f, k, s):
        out.extend(["[convolutional]", "batch_normalize=1", f"filters={f}", f"size={k}", f"stride={s}",
                    "pad=1", "activation=leaky", ""])

    def res(f, n):
        for _ in range(n):
            conv(f // 2, 1, 1)
            conv(f, 3, 1)
            out.extend(["[shortcut]", "from=-3", "activation=linear", ""])

    conv(32, 3, 1)
    conv(64, 3, 2); res(64, 1)
    conv(128, 3, 2); res(128, 2)
    conv(256, 3, 2); res(256, 8)
    conv(512, 3, 2); res(512, 8)
    conv(1024, 3, 2); res(1024, 4)
    conv(512, 1, 1); conv(1024, 3, 1); conv(512, 1, 1); conv(1024, 3, 1); conv(512, 1, 1)
    return "\n".join(out)


def tiny_trunk_cfg():
    """A small cfg with the same block types (stride-1/2 3x3, 1x1, shortcuts) for fast tests."""
    out = ["[net]", "channels=3", "height=64", ""]

    def conv(f, k, s):
        out.extend(["[convolutional]", "batch_normalize=1", f"filters={f}", f"size={k}", f"stride={s}",
                    "pad=1", "activation=leaky", ""])

    conv(32, 3, 1)
    conv(64, 3, 2)
    conv(32, 1, 1); conv(64, 3, 1); out.extend(["[shortcut]", "from=-3", "activation=linear", ""])
    conv(128, 3, 2)
    conv(64, 1, 1); conv(128, 3, 1); out.extend(["[shortcut]", "from=-3", "activation=linear", ""])
    conv(64, 1, 1); conv(128, 3, 1); out.extend(["[shortcut]", "from=-3", "activation=linear", ""])
    conv(64, 1, 1)
    return "\n".join(out)


def parse_cfg_text(text):
    defs = []
    for line in text.split("\n"):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        if line.startswith("["):
            defs.append({"type": line[1:-1].strip()})
            if defs[-1]["type"] == "convolutional":
                defs[-1]["batch_normalize"] = 0
        else:
            k, v = line.split("=")
            defs[-1][k.strip()] = v.strip()
    return defs


def random_trunk_state(cfg_text, seed=0):
    """Random-init state_dict with the reference's names/shapes (nn.Conv2d / nn.BatchNorm2d
    default initialisers, dark_net.py:17-31) for synthetic-weight runs."""
    import torch.nn as nn
    torch.manual_seed(seed)
    defs = parse_cfg_text(cfg_text)
    cin = [int(defs[0]["channels"])]
    sd = {}
    for i, d in enumerate(defs[1:]):
        if d["type"] == "convolutional":
            f, k = int(d["filters"]), int(d["size"])
            sd[f"module_list.{i}.conv_{i}.weight"] = nn.Conv2d(cin[-1], f, k, bias=False).weight.detach().clone()
            sd[f"module_list.{i}.batch_norm_{i}.weight"] = torch.ones(f)
            sd[f"module_list.{i}.batch_norm_{i}.bias"] = torch.zeros(f)
            sd[f"module_list.{i}.batch_norm_{i}.running_mean"] = torch.zeros(f)
            sd[f"module_list.{i}.batch_norm_{i}.running_var"] = torch.ones(f)
            cin.append(f)
        else:
            cin.append(cin[int(d["from"])])
    return sd


def random_et_state(seed=0, d=768, heads=12, layers=2):
    """Random-init ET state_dict (reference parameter names, ET_haa.py:77-119 / enc_vl.py:8-33)
    built from stock torch modules of the same shapes."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}
    enc = nn.TransformerEncoder(nn.TransformerEncoderLayer(d, heads, d, 0.1), layers, enable_nested_tensor=False)
    for k, v in enc.state_dict().items():
        sd["encoder_vl.enc_transformer." + k] = v.detach().clone()
    ln = nn.LayerNorm(d)
    sd["encoder_vl.enc_layernorm.weight"], sd["encoder_vl.enc_layernorm.bias"] = ln.weight.detach().clone(), ln.bias.detach().clone()
    for name, mod in (("decoder_2_action_full.0", nn.Linear(d, 256)), ("decoder_2_action_full.3", nn.Linear(256, 32)),
                      ("decoder_2_action_full.6", nn.Linear(32, 4)), ("fc.0", nn.Linear(d, 64)),
                      ("direction_embedding", nn.Linear(2, d)), ("fc2", nn.Linear(49, d)),
                      ("attention_layer_vision.linear_in", nn.Linear(49, 49, bias=False)),
                      ("attention_layer_vision.linear_out", nn.Linear(98, 49, bias=False))):
        for k, v in mod.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


class _RoundBF16(torch.autograd.Function):
    """Value AND gradient pass through bf16 storage (round-to-nearest-even), fp32 arithmetic
    around it: models a tensor the product keeps in HBM as bf16 (activations, their gradients)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundBF16Fwd(torch.autograd.Function):
    """Forward value rounded to bf16, gradient untouched: models a bf16 copy of an fp32 master
    parameter whose gradient is accumulated in fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def darknet_forward(x, sd, cfg_text, train=True, update_running=False, eps=1e-5, momentum=0.1, storage=None):
    """dark_net.py:212-240 for conv(+BN+leaky) / shortcut blocks.  ``sd`` maps the
    reference state_dict names to tensors (requires_grad leaves for gradient
    parity).  Train mode uses batch statistics (agent.py:214).

    ``storage=None``   the reference arithmetic, float32 end to end.
    ``storage="bf16"`` the SAME arithmetic (float32 products and sums) with every tensor the
    bf16 pipeline stores -- the input image, the conv weights (bf16 copies of the fp32 masters),
    conv outputs, block outputs, and the matching gradients -- rounded to bf16 at the point of
    storage.  This is what "the reference PyTorch
    path in bf16" computes; the random-init 57-block trunk amplifies a perturbation ~100x
    (conditioning), so bf16 parity is checked against this mode and the distance of
    both from the float32 mode is reported beside it."""
    q = _RoundBF16.apply if storage == "bf16" else (lambda t: t)
    qw = _RoundBF16Fwd.apply if storage == "bf16" else (lambda t: t)
    if storage not in (None, "bf16"):
        raise ValueError(storage)
    defs = parse_cfg_text(cfg_text)[1:]
    outs = []
    x = q(x)
    for i, d in enumerate(defs):
        if d["type"] == "convolutional":
            w = qw(sd[f"module_list.{i}.conv_{i}.weight"])
            k = int(d["size"])
            pad = (k - 1) // 2 if int(d["pad"]) else 0
            x = q(F.conv2d(x, w, None, stride=int(d["stride"]), padding=pad))
            g, b = sd[f"module_list.{i}.batch_norm_{i}.weight"], sd[f"module_list.{i}.batch_norm_{i}.bias"]
            rm, rv = sd[f"module_list.{i}.batch_norm_{i}.running_mean"], sd[f"module_list.{i}.batch_norm_{i}.running_var"]
            if train:
                x = F.batch_norm(x, rm if update_running else None, rv if update_running else None, g, b,
                                 True, momentum, eps)
            else:
                x = F.batch_norm(x, rm, rv, g, b, False, momentum, eps)
            x = F.leaky_relu(x, 0.01)                      # nn.LeakyReLU() default slope
            nxt = defs[i + 1] if i + 1 < len(defs) else None
            if not (nxt is not None and nxt["type"] == "shortcut"):
                x = q(x)                                   # a fused shortcut stores only the sum
        elif d["type"] == "shortcut":
            x = q(outs[-1] + outs[int(d["from"])])
        else:
            raise NotImplementedError(d["type"])
        outs.append(x)
    return outs[-1]


# --------------------------------------------------------------------------
# ET (HAA-Transformer)
# --------------------------------------------------------------------------
def soft_dot_attention(h, context, w_in, w_out):
    """ET_haa.py:54-74.  h [B,49]; context [B,512,49]; softmax over the 512 channels."""
    target = h @ w_in.t()                                      # [B,49]
    attn = torch.softmax(torch.bmm(context, target.unsqueeze(2)).squeeze(2), dim=1)   # [B,512]
    weighted = torch.bmm(attn.unsqueeze(1), context).squeeze(1)                      # [B,49]
    return torch.tanh(torch.cat((weighted, h), 1) @ w_out.t()), attn


def pos_encoding_table(d_model=768, max_len=1250):
    """encodings.py:12-20 (float32 table, then /sqrt(d_model) at use)."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def mask_pad(lengths, len_lang):
    """enc_vl.py:44-55: True = padded key.  [B, L + 2*Tmax] bool."""
    tmax = int(np.max(lengths))
    m = torch.zeros((len(lengths), len_lang + 2 * tmax), dtype=torch.bool)
    for i, l in enumerate(lengths):
        m[i, len_lang + l: len_lang + tmax] = True
        m[i, len_lang + tmax + l:] = True
    return m


def attention_mask(len_lang, len_frames):
    """model_util.py:213-241: additive float mask [(L+2T),(L+2T)], 0 = may attend, -inf = may not."""
    S = len_lang + 2 * len_frames
    m = torch.full((S, S), float("-inf"))
    m[:, :len_lang] = 0.0                       # everyone sees language ...
    m[:len_lang, len_lang:] = float("-inf")     # ... language sees only language
    tri = torch.triu(torch.ones(len_frames, len_frames), diagonal=1) == 1      # True above the diagonal
    block = torch.zeros(len_frames, len_frames).masked_fill(tri, float("-inf"))
    for r0 in (len_lang, len_lang + len_frames):
        for c0 in (len_lang, len_lang + len_frames):
            m[r0:r0 + len_frames, c0:c0 + len_frames] = block
    return m


def _layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def transformer_layer(x, sd, prefix, n_heads, mask_attn, mask_padding, drop=None):
    """One post-norm nn.TransformerEncoderLayer(d, h, ff, dropout, relu), batch-first here.
    ``drop`` = None: eval arithmetic (dropout inactive).  Otherwise a dict of explicit keep-scale
    masks (0 or 1/(1-p)) for the layer's four nn.Dropout sites, in torch's order:
    ``attn`` [B,H,S,S] on the softmax output (nn.MultiheadAttention(dropout=p)), ``drop1`` [B,S,E]
    on the attention block's output, ``ffn`` [B,S,FF] after the ReLU, ``drop2`` [B,S,E] on linear2's
    output."""
    B, S, E = x.shape
    dh = E // n_heads
    qkv = x @ sd[prefix + "self_attn.in_proj_weight"].t() + sd[prefix + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(E, dim=-1)
    q = q.view(B, S, n_heads, dh).transpose(1, 2)
    k = k.view(B, S, n_heads, dh).transpose(1, 2)
    v = v.view(B, S, n_heads, dh).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    scores = scores + mask_attn[None, None]
    scores = scores.masked_fill(mask_padding[:, None, None, :], float("-inf"))
    p = torch.softmax(scores, dim=-1)
    if drop is not None:
        p = p * drop["attn"]
    o = (p @ v).transpose(1, 2).reshape(B, S, E)
    o = o @ sd[prefix + "self_attn.out_proj.weight"].t() + sd[prefix + "self_attn.out_proj.bias"]
    if drop is not None:
        o = o * drop["drop1"]
    x = _layer_norm(x + o, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    ff = torch.relu(x @ sd[prefix + "linear1.weight"].t() + sd[prefix + "linear1.bias"])
    if drop is not None:
        ff = ff * drop["ffn"]
    ff = ff @ sd[prefix + "linear2.weight"].t() + sd[prefix + "linear2.bias"]
    if drop is not None:
        ff = ff * drop["drop2"]
    return _layer_norm(x + ff, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])


def encoder_vl_forward(emb_lang, emb_frames, emb_dirs, lengths, sd, n_heads=12, n_layers=2,
                       prefix="encoder_vl.", drop=None):
    """enc_vl.py:34-83.  ``drop``: None or a list (one dict per layer) of explicit dropout masks."""
    B, L, E = emb_lang.shape
    T = emb_frames.shape[1]
    assert T == int(np.max(lengths))
    pe = pos_encoding_table(E)[: L + T] / math.sqrt(E)
    lang = emb_lang + pe[None, :L]
    frames = emb_frames + pe[None, L:L + T]
    dirs = emb_dirs + pe[None, L:L + T]                 # same positions as the frames
    x = torch.cat((lang, frames, dirs), dim=1)
    x = _layer_norm(x, sd[prefix + "enc_layernorm.weight"], sd[prefix + "enc_layernorm.bias"])
    mp = mask_pad(lengths, L)
    ma = attention_mask(L, T)
    for l in range(n_layers):
        x = transformer_layer(x, sd, f"{prefix}enc_transformer.layers.{l}.", n_heads, ma, mp,
                              None if drop is None else drop[l])
    return x, mp


def et_forward(sd, directions, frames, lenths, lang, lang_cls, n_heads=12, n_layers=2, drop=None):
    """ET_haa.py:121-184.  ``drop`` = None: eval arithmetic (dropout inactive); otherwise
    ``{"layers": [per-layer mask dicts], "h0": [B,256], "h1": [B,32], "fc": [B,64]}`` -- explicit
    keep-scale masks of the train-mode nn.Dropout sites (ET_haa.py:98-119: Dropout(0.2) after the two
    hidden ReLUs of decoder_2_action_full and between fc's Linear and its ReLU).
    Returns (output [B,4], pred_saliency [B,1,224,224], h_sali [B,64])."""
    B, T = frames.shape[:2]
    L = lang.shape[1]
    att = []
    for i in range(T):
        a, _ = soft_dot_attention(lang_cls, frames[:, i], sd["attention_layer_vision.linear_in.weight"],
                                  sd["attention_layer_vision.linear_out.weight"])
        att.append(a.unsqueeze(1))
    att = torch.cat(att, dim=1)                                          # [B,T,49]
    emb_frames = att.reshape(-1, 49) @ sd["fc2.weight"].t() + sd["fc2.bias"]
    emb_frames = emb_frames.view(B, T, -1)
    emb_dirs = directions.reshape(-1, 2) @ sd["direction_embedding.weight"].t() + sd["direction_embedding.bias"]
    emb_dirs = emb_dirs.view(B, T, -1)
    enc, _ = encoder_vl_forward(lang, emb_frames, emb_dirs, lenths, sd, n_heads, n_layers,
                                drop=None if drop is None else drop["layers"])
    tmax = int(np.max(lenths))
    vis = enc[:, L + tmax - 1]
    dire = enc[:, L + 2 * tmax - 1]
    h = torch.relu(dire @ sd["decoder_2_action_full.0.weight"].t() + sd["decoder_2_action_full.0.bias"])
    if drop is not None:
        h = h * drop["h0"]
    h = torch.relu(h @ sd["decoder_2_action_full.3.weight"].t() + sd["decoder_2_action_full.3.bias"])
    if drop is not None:
        h = h * drop["h1"]
    output = h @ sd["decoder_2_action_full.6.weight"].t() + sd["decoder_2_action_full.6.bias"]
    hs = vis @ sd["fc.0.weight"].t() + sd["fc.0.bias"]                       # Linear -> Dropout -> ReLU
    if drop is not None:
        hs = hs * drop["fc"]
    h_sali = torch.relu(hs)
    pred = F.interpolate(h_sali.view(-1, 1, 8, 8), size=(224, 224), mode="bilinear", align_corners=False)
    return output, pred, h_sali


# --------------------------------------------------------------------------
# agent slice: loss and waypoint post-processing
# --------------------------------------------------------------------------
def nss(sal, fix, nss_r=0):
    """agent.py:256-270 (unbiased std)."""
    m = torch.mean(sal.view(-1, 224 * 224), 1).view(-1, 1, 1)
    std = torch.std(sal.view(-1, 224 * 224), 1).view(-1, 1, 1)
    n_sal = (sal - m) / std
    if nss_r == 1:
        n_sal = n_sal / 2 + 1
    elif nss_r == -1:
        n_sal = n_sal / 2 - 1
    s_fix = torch.sum(fix.view(-1, 224 * 224), 1) + 0.001
    s_ns = torch.sum((n_sal * fix).view(-1, 224 * 224), 1)
    return -torch.mean(s_ns / s_fix)


def _ang(v0, v1):
    return ((torch.atan2(v0, v1) / PI_REF + 2) / 2) % 1


def et_loss(output, pred_saliency, gt_xy, gt_alt, gt_prog, gt_saliency, nss_w=0.1, nss_r=0, jitter=None):
    """agent.py:663-681 for one time step: sum over ALL samples of the four
    MSE-sum terms + the angular term, plus ``nss_w * NSS`` for samples with
    attention.  ``jitter`` [B] stands for ``1e-5*np.random.rand()`` (agent.py:666;
    zeros = deterministic).  ``gt_saliency`` float64 [B,224,224] -> float64 loss."""
    B = output.shape[0]
    ml = 0
    for i in range(B):
        p_xy = output[i, 0:2]
        ml = ml + torch.sum((p_xy - gt_xy[i]) ** 2)
        j = 0.0 if jitter is None else float(jitter[i])
        ml = ml + (_ang(p_xy[0], p_xy[1] + j) - _ang(gt_xy[i, 0], gt_xy[i, 1])) ** 2
        ml = ml + (output[i, 2] - gt_alt[i]) ** 2
        ml = ml + (output[i, 3] - gt_prog[i]) ** 2
    for i in range(B):
        if float(gt_saliency[i].sum()) > 0:
            v = nss(pred_saliency[i], gt_saliency[i], nss_r)
            if not torch.isnan(v):
                ml = ml + nss_w * v
    return ml


def step_loss(ml_loss, train_ml, batch_size):
    """agent.py:883-885: ``loss += ml_loss * train_ml / batch_size``."""
    return ml_loss * train_ml / batch_size


def postprocess_waypoints(output, edge_len, stop_threshold=0.5):
    """agent.py:637-653,738,745-752 on host float32/float64 exactly as numpy does.
    ``output`` [B,4] float32; ``edge_len`` [B] float64 = ||c0 - c1||.
    Returns (angle_deg int, dist f64, altitude_m int, stop bool) arrays."""
    o = np.asarray(output, dtype=np.float32).copy()
    B = o.shape[0]
    ang = np.zeros(B, dtype=np.int64)
    dist = np.zeros(B, dtype=np.float64)
    alt = np.zeros(B, dtype=np.int64)
    stop = np.zeros(B, dtype=bool)
    for i in range(B):
        x, y = o[i, 0], o[i, 1]
        m = max(abs(x), abs(y), 1)
        x, y = x / m, y / m                                   # float32 divides (numpy array elements)
        a = min(1.0, max(0.0, o[i, 2]))
        p = min(1.0, max(0.0, o[i, 3]))
        a_dir = ((np.arctan2(x, y) / PI_REF + 2) / 2) % 1      # float32 atan2 promoted by the f64 constant
        ang[i] = int(round(float(a_dir) * 360))
        dist[i] = float(np.linalg.norm(np.array([x, y]))) * (float(edge_len[i]) / 2)
        alt[i] = int(round(float(a) * 360)) + 40
        stop[i] = bool(p > stop_threshold)
    return ang, dist, alt, stop


# --------------------------------------------------------------------------
# config 5: ViT_LSTM step and the simulator update of the greedy rollout
# --------------------------------------------------------------------------
def _lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell arithmetic (gate order i, f, g, o); zero state when h is None."""
    gates = x @ w_ih.t() + b_ih + b_hh
    if h is not None:
        gates = gates + h @ w_hh.t()
    H = w_hh.shape[1]
    i, f, g, o = gates[:, :H], gates[:, H:2 * H], gates[:, 2 * H:3 * H], gates[:, 3 * H:]
    c1 = torch.sigmoid(f) * (c if c is not None else 0) + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c1), c1


def vit_lstm_step(sd, im_feature, current_direct, cls_hidden, lang_feature, state=None):
    """src/models/vln_model.py:213-250 after the vision model (eval arithmetic: dropout inactive).
    ``im_feature`` [B,512,49]; ``current_direct`` [B,1] degrees (int64 in the reference loop);
    ``state`` = (h, c, hh, cc) or None.  Returns (h1, c1, hh1, cc1, output [B,4], h_sali [B,64])."""
    h0, c0, hh0, cc0 = state if state is not None else (None, None, None, None)
    inp, _ = soft_dot_attention(cls_hidden, im_feature, sd["attention_layer_vision.linear_in.weight"],
                                sd["attention_layer_vision.linear_out.weight"])
    hh1, cc1 = _lstm_cell(inp, hh0, cc0, sd["vision_lstm.weight_ih"], sd["vision_lstm.weight_hh"],
                          sd["vision_lstm.bias_ih"], sd["vision_lstm.bias_hh"])
    cd = current_direct / 180 * PI_REF                       # int64 / int -> float32 (torch true division)
    direction = torch.cat((torch.sin(cd), torch.cos(cd)), dim=1).float()
    demb = direction @ sd["direction_embedding.weight"].t() + sd["direction_embedding.bias"]
    h1, c1 = _lstm_cell(demb, h0, c0, sd["direct_lstm.weight_ih"], sd["direct_lstm.weight_hh"],
                        sd["direct_lstm.bias_ih"], sd["direct_lstm.bias_hh"])
    # SoftDotAttention(768) over the tokens (vln_model.py:26-46): softmax over seq_len
    hcat = torch.cat((h1, hh1), 1)
    target = hcat @ sd["attention_layer_lang.linear_in.weight"].t()
    attn = torch.softmax(torch.bmm(lang_feature, target.unsqueeze(2)).squeeze(2), dim=1)
    weighted = torch.bmm(attn.unsqueeze(1), lang_feature).squeeze(1)
    act_in = torch.tanh(torch.cat((weighted, hcat), 1) @ sd["attention_layer_lang.linear_out.weight"].t())
    h = torch.relu(act_in @ sd["decoder_2_action_full.0.weight"].t() + sd["decoder_2_action_full.0.bias"])
    h = torch.relu(h @ sd["decoder_2_action_full.3.weight"].t() + sd["decoder_2_action_full.3.bias"])
    output = h @ sd["decoder_2_action_full.6.weight"].t() + sd["decoder_2_action_full.6.bias"]
    s = torch.relu(inp @ sd["fc.0.weight"].t() + sd["fc.0.bias"])
    h_sali = torch.relu(s @ sd["fc.3.weight"].t() + sd["fc.3.bias"])
    return h1, c1, hh1, cc1, output, h_sali


def get_direction(start, end):
    """src/xview_et/agent.py:83-101."""
    vec = np.array(end) - np.array(start)
    if vec[1] > 0:
        ang = np.arctan(vec[0] / vec[1]) / 1.57 * 90
    elif vec[1] < 0:
        ang = np.arctan(vec[0] / vec[1]) / 1.57 * 90 + 180
    else:
        ang = 90 if np.sign(vec[0]) == 1 else 270
    return (360 - ang + 90) % 360


def move_view_corners(corners, angle, distance, altitude, gps_botm_left, gps_top_right, input_current_direction=None):
    """src/xview_et/agent.py:285-384 (identical in src/xview_lstm/agent.py): zoom to the altitude,
    rotate by -angle about the centre (pi = 3.14159), move forward; a stage whose corners leave the map
    is rejected.  Host float64 in the reference's operation order."""
    corners = np.asarray(corners, dtype=np.float64)
    norm = np.linalg.norm

    def inside(p):
        return gps_botm_left[0] < p[0] < gps_top_right[0] and gps_botm_left[1] < p[1] < gps_top_right[1]

    def rot(theta, p):
        M = np.array([[np.cos(theta / 180 * PI_REF), np.sin(theta / 180 * PI_REF)],
                      [-np.sin(theta / 180 * PI_REF), np.cos(theta / 180 * PI_REF)]])
        return np.matmul(M, np.array([p[0], p[1]]))

    def change_corner(cs, ch):
        o = np.zeros((4, 2))
        o[0] = cs[0] + (cs[0] - cs[1]) / norm(cs[1] - cs[0]) * ch
        o[0] += (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
        o[1] = cs[1] + (cs[1] - cs[0]) / norm(cs[1] - cs[0]) * ch
        o[1] += (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
        o[2] = cs[2] + (cs[2] - cs[3]) / norm(cs[2] - cs[3]) * ch
        o[2] += (cs[2] - cs[1]) / norm(cs[2] - cs[1]) * ch
        o[3] = cs[3] + (cs[3] - cs[2]) / norm(cs[2] - cs[3]) * ch
        o[3] += (cs[3] - cs[0]) / norm(cs[3] - cs[0]) * ch
        return o

    def forward(cs, ch):
        o = np.zeros((4, 2))
        o[0] = cs[0] + (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
        o[1] = cs[1] + (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
        o[2] = cs[2] + (cs[1] - cs[2]) / norm(cs[2] - cs[1]) * ch
        o[3] = cs[3] + (cs[0] - cs[3]) / norm(cs[3] - cs[0]) * ch
        return o

    cur = round(get_direction(np.mean(corners, axis=0), (corners[0] + corners[1]) / 2)) % 360
    if input_current_direction is not None and abs(input_current_direction - cur) > 2:
        angle += input_current_direction
    edge = norm(corners[1] - corners[0]) * 11.13 * 1e4
    zoomed = change_corner(corners, 0.5 * (altitude - edge) / 11.13 / 1e4)
    if not all(inside(p) for p in zoomed):
        return np.array(corners), cur
    centre = np.mean(zoomed, axis=0)
    rotated = [centre + rot(-angle, zoomed[i] - centre) for i in range(4)]
    if not all(inside(p) for p in rotated):
        return np.array(zoomed), cur
    moved = forward(np.array(rotated), distance)
    if not all(inside(p) for p in moved):
        return np.array(rotated), (cur + angle) % 360
    return np.array(moved), (cur + angle) % 360


def waypoint_step(output, corners, bounds, cur_dir, ended, stop_threshold=0.25, last_step=False):
    """One simulator step of the student rollout (src/xview_lstm/agent.py:607-626,700-730) for a batch.
    ``output`` [B,4] f32; ``corners`` [B,4,2] f64; ``bounds`` [B,4] = (bl_lat, bl_lng, tr_lat, tr_lng);
    ``cur_dir`` [B]; ``ended`` [B] bool.  Returns new (corners, cur_dir, ended, angle, altitude, dist)."""
    import math
    o = np.asarray(output, dtype=np.float32).copy()
    B = o.shape[0]
    corners = np.array(corners, dtype=np.float64)
    cur_dir = np.array(cur_dir, dtype=np.float64)
    ended = np.array(ended, dtype=bool)
    ang = np.zeros(B, dtype=np.int64)
    alt = np.zeros(B, dtype=np.int64)
    dist = np.zeros(B, dtype=np.float64)
    for i in range(B):
        m = max(abs(o[i, 0]), abs(o[i, 1]), 1)
        o[i, 0] /= m
        o[i, 1] /= m
        a = min(1., max(0., o[i, 2]))
        p = min(1., max(0., o[i, 3]))
        a_dir = (math.atan2(o[i, 0], o[i, 1]) / PI_REF + 2) / 2 % 1
        ang[i] = round(a_dir * 360)
        dist[i] = np.linalg.norm(o[i, 0:2]) * (np.linalg.norm(corners[i][0] - corners[i][1]) / 2)
        alt[i] = round(a * 360) + 40
        if p > stop_threshold or last_step:
            ended[i] = True
            continue
        new_c, new_d = move_view_corners(corners[i], int(ang[i]), dist[i], int(alt[i]), bounds[i, 0:2], bounds[i, 2:4],
                                         cur_dir[i])
        corners[i], cur_dir[i] = new_c, new_d
    return corners, cur_dir, ended, ang, alt, dist
