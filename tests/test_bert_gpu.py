"""GPU: the language encoder (CustomBERTModel on libavdn kernels) against the oracle that is
pinned to transformers.BertModel (tests/test_bert_oracle.py).  bf16 tensor-core path: 1e-2 on outputs, 5e-2
relative L2 on every parameter gradient."""
import pytest
import torch

from oracle import bert_oracle as bo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-6)


def _rel2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _model(layers, vocab, seed):
    from transformers import BertConfig
    from avdn_b200.models.bert import CustomBERTModel
    torch.manual_seed(seed)
    m = CustomBERTModel(BertConfig(num_hidden_layers=layers, vocab_size=vocab)).cuda()
    # HF initialises biases / LayerNorm to constants: perturb them so that every gradient path is exercised
    g = torch.Generator(device="cuda").manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias") or "LayerNorm" in n:
                p.add_(torch.randn(p.shape, device="cuda", generator=g) * 0.05)
    return m.eval()                               # deterministic arithmetic; the dropout test switches to train()


def _case(B, S, seed, vocab):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, S), generator=g)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    lens[0] = S
    mask = (torch.arange(S)[None] < lens[:, None]).long()
    return ids, mask


def test_state_dict_keys_are_the_reference_modules(built_lib):
    m = _model(1, 100, 0)
    keys = set(m.state_dict())
    assert "bert.embeddings.word_embeddings.weight" in keys and "bert.encoder.layer.0.attention.self.query.weight" in keys
    assert "bert.pooler.dense.weight" in keys and "linears.0.weight" in keys and "linears.3.bias" in keys


def test_forward_and_all_gradients_vs_oracle(built_lib):
    B, S, V = 3, 40, 1000
    m = _model(3, V, 2)
    ids, mask = _case(B, S, 3, V)
    seq, lin, cls = m(ids.cuda(), mask.cuda())
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()
          if "position_ids" not in k}
    seq_o, lin_o, cls_o = bo.custom_bert_forward(sd, ids, mask)
    assert _rel(seq, seq_o) < 1e-2, _rel(seq, seq_o)
    assert _rel(cls, cls_o) < 1e-2 and _rel(lin, lin_o) < 2e-2, (_rel(cls, cls_o), _rel(lin, lin_o))
    g = torch.Generator().manual_seed(5)
    w_seq = torch.randn(B, S, 768, generator=g) * mask[..., None]
    w_lin, w_cls = torch.randn(B, 49, generator=g), torch.randn(B, 768, generator=g)
    ((seq_o * w_seq).sum() + (lin_o * w_lin).sum() + (cls_o * w_cls).sum()).backward()
    m.zero_grad()
    ((seq * w_seq.cuda()).sum() + (lin * w_lin.cuda()).sum() + (cls * w_cls.cuda()).sum()).backward()
    worst = {}
    for n, p in m.named_parameters():
        ref = sd[n].grad
        assert p.grad is not None and ref is not None, n
        if ref.norm() < 1e-4:                     # key biases: the gradient is identically zero
            assert p.grad.float().cpu().norm() < 5e-2, n
            continue
        worst[n] = _rel2(p.grad, ref)
        assert worst[n] < 5e-2, (n, worst[n])


def _site_mask(n, p, seed, site):
    from avdn_b200 import _lib
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _lib.call("avdn_dropout_keep_scale", _lib.ptr(out), n, float(p), int(seed), int(site))
    return out.cpu()


def test_train_mode_dropout_vs_oracle_with_same_masks(built_lib):
    """Train mode applies HF's dropout sites and the head's Dropout(0.2) with stateless hash masks; the step's
    masks are read back and fed to the oracle: outputs and every gradient must agree."""
    B, S, V, NL, H = 3, 40, 1000, 2, 12
    Sp = 64
    m = _model(NL, V, 4)
    ids, mask = _case(B, S, 5, V)
    m.train()
    try:
        seq, lin, cls = m(ids.cuda(), mask.cuda())
        eng = m.engine(B, S, seq.device)
        ph, pa, pd, seed = eng.p_hid, eng.p_att, eng.p_head, eng.seed
        assert (ph, pa, pd) == (pytest.approx(0.1), pytest.approx(0.1), pytest.approx(0.2)) and eng.Sp == Sp
        g = torch.Generator().manual_seed(6)
        w_seq = torch.randn(B, S, 768, generator=g) * mask[..., None]
        w_lin, w_cls = torch.randn(B, 49, generator=g), torch.randn(B, 768, generator=g)
        m.zero_grad()
        ((seq * w_seq.cuda()).sum() + (lin * w_lin.cuda()).sum() + (cls * w_cls.cuda()).sum()).backward()
    finally:
        m.eval()
    layers = []
    for l in range(NL):
        st = eng.SITE_LAYER0 + 3 * l
        layers.append(dict(attn=_site_mask(B * H * S * Sp, pa, seed, st).view(B, H, S, Sp)[..., :S].contiguous(),
                           ao=_site_mask(B * S * 768, ph, seed, st + 1).view(B, S, 768),
                           fo=_site_mask(B * S * 768, ph, seed, st + 2).view(B, S, 768)))
    drop = dict(emb=_site_mask(B * S * 768, ph, seed, eng.SITE_EMB).view(B, S, 768), layers=layers,
                head=_site_mask(B * 64, pd, seed, eng.SITE_HEAD).view(B, 64))
    keep = (drop["emb"] > 0).float().mean().item()
    assert abs(keep - 0.9) < 5e-3, keep
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()
          if "position_ids" not in k}
    seq_o, lin_o, cls_o = bo.custom_bert_forward(sd, ids, mask, drop=drop)
    assert _rel(seq, seq_o) < 1e-2, _rel(seq, seq_o)
    assert _rel(cls, cls_o) < 1e-2 and _rel(lin, lin_o) < 2e-2
    ((seq_o * w_seq).sum() + (lin_o * w_lin).sum() + (cls_o * w_cls).sum()).backward()
    for n, p in m.named_parameters():
        ref = sd[n].grad
        if ref.norm() < 1e-4:
            assert p.grad.float().cpu().norm() < 5e-2, n
            continue
        assert _rel2(p.grad, ref) < 5e-2, (n, _rel2(p.grad, ref))


def test_full_depth_forward(built_lib):
    """12 layers, the real vocabulary, B=4 x S=64 with ragged padding."""
    B, S, V = 4, 64, 30522
    m = _model(12, V, 7)
    ids, mask = _case(B, S, 8, V)
    with torch.no_grad():
        seq, lin, cls = m(ids.cuda(), mask.cuda())
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if "position_ids" not in k}
        seq_o, lin_o, cls_o = bo.custom_bert_forward(sd, ids, mask)
    assert seq.shape == (B, S, 768) and lin.shape == (B, 49) and cls.shape == (B, 768)
    assert _rel(seq, seq_o) < 2e-2, _rel(seq, seq_o)
    assert _rel(cls, cls_o) < 2e-2, _rel(cls, cls_o)
