"""GPU: edge shapes of the hot path -- smallest batches / sequences, odd sizes -- against the oracle."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import bert_oracle as bo
from oracle import model_oracle as mo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-6)


def test_et_single_sample_single_step(built_lib):
    from avdn_b200.models.ET_haa import ET
    import types
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0)
    torch.manual_seed(3)
    et = ET(args).cuda().eval()
    for (B, L, T, lens) in ((1, 1, 1, [1]), (1, 7, 3, [3]), (2, 250, 1, [1, 1]), (5, 33, 20, [20, 1, 7, 19, 2])):
        g = torch.Generator().manual_seed(B * 100 + T)
        lang = torch.randn(B, L, 768, generator=g)
        cls = torch.relu(torch.randn(B, 49, generator=g))
        frames = torch.randn(B, T, 512, 49, generator=g) * 0.5
        dirs = torch.randn(B, T, 2, generator=g)
        sd = {k: v.detach().cpu() for k, v in et.state_dict().items()}
        with torch.no_grad():
            oo, sal_o, _ = mo.et_forward(sd, dirs, frames, lens, lang, cls)
            out, sal = et(directions=dirs.cuda(), frames=frames.cuda(), lenths=lens, lang=lang.cuda(), lang_cls=cls.cuda())
        assert out.shape == (B, 4) and sal.shape == (B, 1, 224, 224)
        assert _rel(out, oo) < 1e-2, (B, L, T, _rel(out, oo))
        assert _rel(sal, sal_o) < 1e-2, (B, L, T)


def test_trunk_single_image_eval_and_train(built_lib):
    from avdn_b200.models.dark_net import Darknet
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.tiny_trunk_cfg())
    torch.manual_seed(0)
    net = Darknet(f.name, 64).cuda()
    os.unlink(f.name)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    x = torch.randn(1, 3, 64, 64)
    net.eval()
    with torch.no_grad():
        y = net(x.cuda())
    ref = mo.darknet_forward(x, sd, mo.tiny_trunk_cfg(), train=False)
    assert y.shape == ref.shape and _rel(y, ref) < 1e-2
    net.train()
    y2 = net(x.cuda())                                   # batch statistics over one image
    ref2 = mo.darknet_forward(x, sd, mo.tiny_trunk_cfg(), train=True)
    assert _rel(y2, ref2) < 3e-2, _rel(y2, ref2)
    y2.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_bert_shortest_sequences(built_lib):
    from transformers import BertConfig
    from avdn_b200.models.bert import CustomBERTModel
    torch.manual_seed(1)
    m = CustomBERTModel(BertConfig(num_hidden_layers=2, vocab_size=300)).cuda().eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items() if "position_ids" not in k}
    for (B, S) in ((1, 1), (1, 2), (3, 5), (2, 65)):
        g = torch.Generator().manual_seed(S)
        ids = torch.randint(0, 300, (B, S), generator=g)
        mask = torch.ones(B, S, dtype=torch.long)
        if B > 1 and S > 2:
            mask[1, S // 2:] = 0
        with torch.no_grad():
            seq, lin, cls = m(ids.cuda(), mask.cuda())
            seq_o, lin_o, cls_o = bo.custom_bert_forward(sd, ids, mask)
        assert _rel(seq, seq_o) < 1e-2, (B, S, _rel(seq, seq_o))
        assert _rel(cls, cls_o) < 1e-2 and _rel(lin, lin_o) < 2e-2


def test_render_single_pose_and_tiny_tile(built_lib):
    from avdn_b200.env import ViewRenderer
    r = ViewRenderer("cuda")
    tile = wo.synthetic_tile(seed=9, size=40)            # smaller than one view: heavy up-sampling + borders
    r.add_map("tiny", tile, None)
    c = np.array([[[5, 6], [30, 9], [27, 33], [2, 30]]], dtype=np.int32)
    v = r.render(torch.from_numpy(c), torch.zeros(1, dtype=torch.int32))["views"].cpu().numpy()
    assert np.array_equal(v[0], wo.render_view(tile, c[0]))
    c2 = np.array([[[-30, -30], [80, -25], [75, 90], [-35, 85]]], dtype=np.int32)   # footprint larger than the tile
    v2 = r.render(torch.from_numpy(c2), torch.zeros(1, dtype=torch.int32))["views"].cpu().numpy()
    assert np.array_equal(v2[0], wo.render_view(tile, c2[0]))
