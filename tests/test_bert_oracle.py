"""CPU: the BERT restatement (oracle/bert_oracle.py) pinned against the third-party implementation the
reference calls (transformers.BertModel behind CustomBERTModel, src/models/vln_model.py:128-159)."""
import torch

from oracle import bert_oracle as bo


def _case(B, S, seed, vocab):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, S), generator=g)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    lens[0] = S
    mask = (torch.arange(S)[None] < lens[:, None]).long()
    return ids, mask


def test_bert_oracle_matches_hf_forward_and_gradients():
    sd, hf = bo.random_custom_bert_state(seed=0, layers=3, vocab=1000)
    hf.eval()
    ids, mask = _case(3, 21, 1, 1000)
    ref = hf(ids, attention_mask=mask)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    seq, pooled = bo.bert_forward(sdg, ids, mask)
    assert torch.allclose(seq, ref["last_hidden_state"], atol=2e-5, rtol=1e-5)
    assert torch.allclose(pooled, ref["pooler_output"], atol=2e-5, rtol=1e-5)
    w = torch.randn_like(seq) * mask[..., None]
    wp = torch.randn_like(pooled)
    ((seq * w).sum() + (pooled * wp).sum()).backward()
    hf.zero_grad()
    ((ref["last_hidden_state"] * w).sum() + (ref["pooler_output"] * wp).sum()).backward()
    for k, p in hf.named_parameters():
        g = sdg["bert." + k].grad
        assert g is not None, k
        assert (g - p.grad).norm() <= 1e-4 * p.grad.norm() + 2e-5, k      # (key-bias gradients are identically 0)


def test_custom_bert_head_shapes():
    sd, _ = bo.random_custom_bert_state(seed=1, layers=1, vocab=500)
    ids, mask = _case(2, 9, 2, 500)
    seq, lin, cls = bo.custom_bert_forward(sd, ids, mask)
    assert seq.shape == (2, 9, 768) and lin.shape == (2, 49) and cls.shape == (2, 768)
    assert (lin >= 0).all()
