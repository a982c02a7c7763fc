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
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-6)


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
    assert _rel(y, g["y"]) < 1e-2, _rel(y, g["y"])
    y.backward(g["dy"].cuda())
    for n, p in net.named_parameters():
        if n in g["grads"]:
            r = _rel(p.grad, g["grads"][n])
            assert r < 5e-2, (n, r)
    # all gradients against the oracle
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in g["sd"].items()}
    yo = mo.darknet_forward(g["x"], sd, g["cfg"], train=True)
    yo.backward(g["dy"])
    for n, p in net.named_parameters():
        r = _rel(p.grad, sd[n].grad)
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


def test_full_trunk_shapes_224(built_lib):
    """The real truncated yolov3 trunk at 224x224: 80 modules, 51,602,144 params, [N,512,7,7]."""
    from avdn_b200.models.dark_net import Darknet
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    torch.manual_seed(0)
    net = Darknet(f.name, 224).cuda()
    os.unlink(f.name)
    assert sum(p.numel() for p in net.parameters()) == 51602144
    net.train()
    x = torch.randn(2, 3, 224, 224, device="cuda")
    y = net(x)
    assert y.shape == (2, 512, 7, 7)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ref = mo.darknet_forward(x.cpu(), sd, mo.yolov3_trunk_cfg(), train=True)
    # 57 train-mode BN layers at batch 2 amplify rounding; statistics-level agreement
    assert _rel(y, ref) < 0.15
    y.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
