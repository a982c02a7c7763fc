"""GPU: supervision geometry (compute_iou + teacher_action, student feedback) against the
oracle (pinned to OpenCV / bisection in tests/test_teacher_oracle.py) on GPS-scale coordinates."""
import numpy as np
import pytest
import torch

from oracle import teacher_oracle as to

pytestmark = pytest.mark.gpu


def _quad(rng, c, side, th=None):
    th = rng.uniform(0, 2 * np.pi) if th is None else th
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return c + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * side / 2) @ R.T


@pytest.mark.parametrize("feedback", ["student", "teacher"])
def test_teacher_action_kernel_vs_oracle(built_lib, feedback):
    from avdn_b200 import _lib
    rng = np.random.default_rng(0)
    B, pmax = 512, 7
    base = np.array([40.01, -74.99])
    corners = np.zeros((B, 4, 2))
    gt = np.zeros((B, pmax, 4, 2))
    lens = rng.integers(1, pmax + 1, size=B)
    ended = rng.random(B) < 0.15
    for i in range(B):
        c = base + rng.uniform(-0.005, 0.005, 2)
        corners[i] = _quad(rng, c, rng.uniform(0.0008, 0.004))
        path_c = c + rng.uniform(-0.006, 0.006, 2)
        for j in range(lens[i]):
            path_c = path_c + rng.uniform(-0.002, 0.002, 2)
            gt[i, j] = _quad(rng, path_c, rng.uniform(0.0008, 0.004))
        if i % 7 == 0:                                   # the view sits (almost) on the goal: progress > 0.5
            gt[i, lens[i] - 1] = corners[i] + rng.uniform(-1e-4, 1e-4, 2)
        if i % 11 == 0:                                  # the goal centre is inside the view
            gt[i, lens[i] - 1] = _quad(rng, corners[i].mean(0) + rng.uniform(-2e-4, 2e-4, 2), 0.0005)
    d = "cuda"
    ratio = torch.empty((B, 2), dtype=torch.float32, device=d)
    alt = torch.empty(B, dtype=torch.float32, device=d)
    prog = torch.empty(B, dtype=torch.float32, device=d)
    ptr = _lib.ptr
    args = (torch.from_numpy(corners).to(d), torch.from_numpy(gt).to(d), torch.from_numpy(lens.astype(np.int32)).to(d),
            torch.from_numpy(ended.astype(np.uint8)).to(d))
    _lib.call("avdn_teacher_action", ptr(args[0]), ptr(args[1]), pmax, ptr(args[2]), ptr(args[3]), B,
              int(feedback == "teacher"), ptr(ratio), ptr(alt), ptr(prog))
    ratio, alt, prog = ratio.cpu().numpy(), alt.cpu().numpy(), prog.cpu().numpy()
    n_zero = n_inside = 0
    for i in range(B):
        r, a, p = to.teacher_action(corners[i], gt[i, :lens[i]], bool(ended[i]), feedback=feedback)
        assert abs(prog[i] - p) <= 1e-6 * max(1.0, abs(p)) + 1e-7, (i, prog[i], p)
        assert abs(alt[i] - a) <= 1e-5 * max(1.0, abs(a)), (i, alt[i], a)
        np.testing.assert_allclose(ratio[i], r, rtol=1e-5, atol=1e-6, err_msg=str(i))
        n_zero += bool(np.all(r == 0))
        n_inside += bool(np.max(np.abs(r)) < 0.98 and not np.all(r == 0))
    assert n_zero > 50 and n_inside > 10                  # ended / arrived / goal-inside-view branches were hit


def test_agent_teacher_action_api(built_lib):
    import os, tempfile, types
    from oracle import model_oracle as mo
    from avdn_b200.xview_et.agent import NavCMTAgent
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.tiny_trunk_cfg())
    args = types.SimpleNamespace(demb=768, encoder_heads=12, encoder_layers=2, dropout_transformer_encoder=0.1,
                                 num_input_actions=1, dropout_emb=0.0, darknet_model_file=f.name, darknet_weight_file=None)
    agent = NavCMTAgent(args, device="cuda")
    os.unlink(f.name)
    rng = np.random.default_rng(3)
    base = np.array([40.01, -74.99])
    corners = [_quad(rng, base, 0.003) for _ in range(3)]
    paths = [[_quad(rng, base + rng.uniform(-0.004, 0.004, 2), 0.002) for _ in range(n)] for n in (1, 4, 2)]
    ratio, alt, prog = agent.teacher_action(corners, paths, [False, False, True], feedback="teacher")
    for i in range(3):
        r, a, p = to.teacher_action(corners[i], paths[i], i == 2, feedback="teacher")
        np.testing.assert_allclose(ratio[i].cpu().numpy(), r, rtol=1e-5, atol=1e-6)
        assert abs(float(alt[i]) - a) < 1e-4 and abs(float(prog[i]) - p) < 1e-6
