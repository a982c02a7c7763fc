"""GPU: the CUDA Darknet trunk (tcgen05 convs + fused BN kernels, bf16) against the
fp32 oracle on identical weights.  Tolerance: 1e-2 relative for the bf16 path
(north_star); gradients of a 9-conv bf16 chain are checked at 5e-2 of each
tensor's max magnitude."""
import os
import tempfile

import pytest
import torch

from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu


def _rel(a, b):
    """max-abs error relative to the largest reference magnitude"""
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-6)


def _rel2(a, b):
    """relative L2 error"""
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def tiny(built_lib, golden_dir):
    from avdn_b200.models.dark_net import Darknet
    g = torch.load(os.path.join(golden_dir, "model_golden.pt"), weights_only=False)["darknet"]
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(g["cfg"])
    net = Darknet(f.name, 64).cuda()
    os.unlink(f.name)
    missing = net.load_state_dict(g["sd"], strict=True)
    return net, g


def test_state_dict_keys_match_reference(tiny):
    net, g = tiny
    assert sorted(net.state_dict().keys()) == sorted(g["sd"].keys())


def test_train_forward_backward_vs_golden_and_oracle(tiny):
    net, g = tiny
    net.train()
    net.load_state_dict(g["sd"])
    x = g["x"].cuda()
    y = net(x)
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    # 10 chained bf16 conv+BN blocks: each block is within 1e-2 (see the layer-wise test below);
    # the chain accumulates ~sqrt(10) x 0.5 % of bf16 rounding
    assert _rel2(y, g["y"]) < 3e-2, _rel2(y, g["y"])
    assert _rel(y, g["y"]) < 3e-2, _rel(y, g["y"])
    y.backward(g["dy"].cuda())
    for n, p in net.named_parameters():
        if n in g["grads"]:
            r = _rel2(p.grad, g["grads"][n])
            assert r < 5e-2, (n, r)
    # all gradients against the oracle
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in g["sd"].items()}
    yo = mo.darknet_forward(g["x"], sd, g["cfg"], train=True)
    yo.backward(g["dy"])
    for n, p in net.named_parameters():
        r = _rel2(p.grad, sd[n].grad)
        assert r < 5e-2, (n, r)
    for k, v in g["running_after"].items():
        assert _rel(net.state_dict()[k], v) < 1e-2, k


def test_eval_forward(tiny):
    net, g = tiny
    net.load_state_dict(g["sd"])
    net.eval()
    with torch.no_grad():
        y = net(g["x"].cuda())
    ref = mo.darknet_forward(g["x"], g["sd"], g["cfg"], train=False)
    assert _rel(y, ref) < 1e-2


def test_second_step_reuses_plans_and_accumulates(tiny):
    net, g = tiny
    net.train()
    net.load_state_dict(g["sd"])
    net.zero_grad()
    x = g["x"].cuda()
    net(x).backward(g["dy"].cuda())
    g1 = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.load_state_dict(g["sd"])
    net(x).backward(g["dy"].cuda())
    for n, p in net.named_parameters():
        assert _rel(p.grad, 2 * g1[n]) < 1e-3, n


def test_full_trunk_224_layerwise(built_lib):
    """The real truncated yolov3 trunk at 224x224: 80 modules, 51,602,144 params, [N,512,7,7].
    Every one of the 57 conv blocks is checked in isolation: the oracle block is fed OUR
    input activation of that block (teacher forcing), so each comparison sees one layer of
    bf16 rounding only."""
    import torch.nn.functional as F
    from avdn_b200.models.dark_net import Darknet
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    torch.manual_seed(0)
    net = Darknet(f.name, 224).cuda()
    os.unlink(f.name)
    assert sum(p.numel() for p in net.parameters()) == 51602144
    assert len(net.module_list) == 80
    net.train()
    N = 4
    x = torch.randn(N, 3, 224, 224, device="cuda")
    y = net(x)
    assert y.shape == (N, 512, 7, 7)
    eng = list(net._engines.values())[0]
    assert len(eng.layers) == 57
    worst = 0.0
    for L in eng.layers:
        conv, bn = net.module_list[L.idx][0], net.module_list[L.idx][1]
        if L.first:
            xin = eng.x_in[..., :3].float().permute(0, 3, 1, 2)
        else:
            xin = L.src.a[..., :L.Cin].float().permute(0, 3, 1, 2)
        z = F.conv2d(xin, conv.weight.float(), None, stride=L.s, padding=(L.k - 1) // 2)
        a = F.leaky_relu(F.batch_norm(z, None, None, bn.weight, bn.bias, True, 0.1, 1e-5), 0.01)
        if L.res is not None:
            a = a + L.res.a[..., :L.Cout].float().permute(0, 3, 1, 2)
        ours = L.a[..., :L.Cout].float().permute(0, 3, 1, 2)
        r = _rel2(ours, a)
        worst = max(worst, r)
        assert r < 1e-2, (L.idx, L.Cin, L.Cout, L.k, L.s, r)
        if L.Cout_p > L.Cout:
            assert (L.a[..., L.Cout:] == 0).all()
    y.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    # last-block gradients against autograd of the same block on our activations
    L = eng.layers[-1]
    conv, bn = net.module_list[L.idx][0], net.module_list[L.idx][1]
    xin = L.src.a[..., :L.Cin].float().permute(0, 3, 1, 2)
    w = conv.weight.detach().clone().requires_grad_(True)
    gm = bn.weight.detach().clone().requires_grad_(True)
    bt = bn.bias.detach().clone().requires_grad_(True)
    a = F.leaky_relu(F.batch_norm(F.conv2d(xin, w, None, stride=L.s, padding=(L.k - 1) // 2), None, None, gm, bt,
                                  True, 0.1, 1e-5), 0.01)
    a.square().mean().backward()
    assert _rel2(conv.weight.grad, w.grad) < 3e-2
    assert _rel2(bn.weight.grad, gm.grad) < 3e-2 and _rel2(bn.bias.grad, bt.grad) < 3e-2
