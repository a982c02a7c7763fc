"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the language encoder that feeds the AVDN hot path
(SURVEY.md §8f N1): ``CustomBERTModel`` (src/models/vln_model.py:128-159) = HuggingFace ``BertModel``
('bert-base-uncased' architecture; the reference pins transformers==4.19.4, requirements.txt:10) followed by
``linears`` = Linear(768,64) -> ReLU -> Dropout(0.2) -> Linear(64,49) -> ReLU on the pooler output.

The BERT arithmetic lives in the third-party library, so it is restated here from the published architecture
(Devlin et al. 2018; post-LayerNorm encoder, erf-GELU, LayerNorm eps 1e-12, learned position embeddings, additive
key-padding mask) in plain torch ops and pinned against ``transformers.BertModel`` itself
(tests/test_bert_oracle.py: random-init weights, eval mode, forward 1e-5 / gradients 1e-4).  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import this module.

State-dict keys are the reference module's: ``bert.<hf key>`` and ``linears.{0,3}.{weight,bias}``.
"""
import math

import torch
import torch.nn.functional as F


def bert_forward(sd, ids, mask, n_heads=12, eps=1e-12, prefix="bert.", drop=None):
    """ids [B,S] int64, mask [B,S] (1 = token, 0 = padding).  Returns (sequence_output [B,S,768],
    pooler_output [B,768]).  ``drop`` = None: eval arithmetic (dropout inactive); otherwise explicit keep-scale
    masks of HF's train-mode dropout sites: ``{"emb": [B,S,E], "layers": [{"attn": [B,H,S,S], "ao": [B,S,E],
    "fo": [B,S,E]}, ...]}`` (BertEmbeddings.dropout, BertSelfAttention.dropout on the probabilities,
    BertSelfOutput.dropout, BertOutput.dropout)."""
    B, S = ids.shape
    p = lambda k: sd[prefix + k]
    E = p("embeddings.word_embeddings.weight").shape[1]
    x = (p("embeddings.word_embeddings.weight")[ids] + p("embeddings.position_embeddings.weight")[:S][None]
         + p("embeddings.token_type_embeddings.weight")[0][None, None])
    x = F.layer_norm(x, (E,), p("embeddings.LayerNorm.weight"), p("embeddings.LayerNorm.bias"), eps)
    if drop is not None:
        x = x * drop["emb"]
    dh = E // n_heads
    neg = torch.finfo(x.dtype).min
    add_mask = (1.0 - mask.to(x.dtype))[:, None, None, :] * neg          # HF: (1 - mask) * finfo.min on the keys
    l = 0
    while (prefix + f"encoder.layer.{l}.attention.self.query.weight") in sd:
        pre = f"encoder.layer.{l}."
        lin = lambda name, t: t @ p(pre + name + ".weight").t() + p(pre + name + ".bias")
        q = lin("attention.self.query", x).view(B, S, n_heads, dh).transpose(1, 2)
        k = lin("attention.self.key", x).view(B, S, n_heads, dh).transpose(1, 2)
        v = lin("attention.self.value", x).view(B, S, n_heads, dh).transpose(1, 2)
        sc = q @ k.transpose(-1, -2) / math.sqrt(dh) + add_mask
        pr = torch.softmax(sc, dim=-1)
        if drop is not None:
            pr = pr * drop["layers"][l]["attn"]
        ctx = (pr @ v).transpose(1, 2).reshape(B, S, E)
        a = lin("attention.output.dense", ctx)
        if drop is not None:
            a = a * drop["layers"][l]["ao"]
        x = F.layer_norm(x + a, (E,), p(pre + "attention.output.LayerNorm.weight"),
                         p(pre + "attention.output.LayerNorm.bias"), eps)
        u = lin("intermediate.dense", x)
        h = 0.5 * u * (1.0 + torch.erf(u / math.sqrt(2.0)))                # hidden_act = "gelu" (erf form)
        o = lin("output.dense", h)
        if drop is not None:
            o = o * drop["layers"][l]["fo"]
        x = F.layer_norm(x + o, (E,), p(pre + "output.LayerNorm.weight"), p(pre + "output.LayerNorm.bias"), eps)
        l += 1
    pooled = torch.tanh(x[:, 0] @ p("pooler.dense.weight").t() + p("pooler.dense.bias"))
    return x, pooled


def custom_bert_forward(sd, ids, mask, drop=None):
    """CustomBERTModel.forward (vln_model.py:149-159): (sequence_output, linear_output [B,49], cls_hidden).
    ``drop["head"]`` [B,64]: the keep-scale mask of ``linears[2]`` = Dropout(0.2)."""
    seq, pooled = bert_forward(sd, ids, mask, drop=drop)
    h = torch.relu(pooled @ sd["linears.0.weight"].t() + sd["linears.0.bias"])
    if drop is not None:
        h = h * drop["head"]
    lin = torch.relu(h @ sd["linears.3.weight"].t() + sd["linears.3.bias"])
    return seq, lin, pooled


def random_custom_bert_state(seed=0, layers=12, vocab=30522):
    """Random-init state dict with the reference module's keys (HF BertModel's default init under ``seed``)."""
    from transformers import BertConfig, BertModel
    import torch.nn as nn
    torch.manual_seed(seed)
    cfg = BertConfig(num_hidden_layers=layers, vocab_size=vocab)
    m = BertModel(cfg)
    sd = {"bert." + k: v.detach().clone() for k, v in m.state_dict().items() if "position_ids" not in k}
    lin = nn.Sequential(nn.Linear(768, 64), nn.ReLU(), nn.Dropout(0.2), nn.Linear(64, 49), nn.ReLU())
    for k, v in lin.state_dict().items():
        sd["linears." + k] = v.detach().clone()
    return sd, m
