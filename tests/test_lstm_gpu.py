"""GPU: config 5 -- the CUDA ViT_LSTM step, the simulator-update kernel and the on-device greedy rollout --
against the reference's golden vectors and the oracle.

Tolerances: the policy step is fp32 on CUDA cores -> 1e-4 relative vs the reference golden; discretised
angle / altitude / stop flags / headings bit-exact; GPS corners 1e-12 relative (float64, same operation
order; device libm vs numpy differ in the last ulp of sin / cos / atan)."""
import os
import tempfile
import types

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden(built_lib, golden_dir):
    return torch.load(os.path.join(golden_dir, "lstm_golden.pt"), weights_only=False)


def _model(g):
    from avdn_b200.models.vln_model import ViT_LSTM

    class Ident(torch.nn.Module):
        def forward(self, x):
            return x

    torch.manual_seed(3)
    m = ViT_LSTM(types.SimpleNamespace(), Ident())
    sd = m.state_dict()
    assert sorted(sd.keys()) == g["keys"]                 # same state_dict keys as the reference module
    for k, v in g["sd_small"].items():
        assert torch.equal(sd[k], v), k
    return m.cuda().eval()


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def test_vit_lstm_forward_two_steps_vs_reference_golden(golden):
    g = golden["lstm"]
    m = _model(g)
    c = lambda t: t.cuda()
    h, cc_, hh, ccc, o1, s1 = m(c(g["d1"]), c(g["feat1"]), None, c(g["cls_hidden"]), c(g["lang"]))
    assert _rel(o1, g["out1"]) < 1e-4, _rel(o1, g["out1"])
    h2, c2, hh2, cc2, o2, s2 = m(c(g["d2"]), c(g["feat2"]), None, c(g["cls_hidden"]), c(g["lang"]), h, cc_, hh, ccc)
    assert _rel(o2, g["out2"]) < 1e-4, _rel(o2, g["out2"])
    for a, k in zip((h2, c2, hh2, cc2), ("h2", "c2", "hh2", "cc2")):
        assert _rel(a, g[k]) < 1e-4, k
    assert s2.shape == (3, 1, 224, 224)
    assert _rel(s2[:, :, ::16, ::16], g["sal2_sub"]) < 1e-4


def test_train_mode_is_refused(golden):
    m = _model(golden["lstm"])
    m.train()
    with pytest.raises(NotImplementedError):
        m.step(torch.zeros(1, 512, 49).cuda(), torch.zeros(1).cuda(), torch.zeros(1, 49).cuda(),
               torch.zeros(1, 4, 768).cuda())


def _poses(B, seed):
    rng = np.random.default_rng(seed)
    bl, tr = np.array([40.0, -75.0]), np.array([40.02, -74.98])
    corners = np.zeros((B, 4, 2))
    dirs = np.zeros(B)
    for i in range(B):
        ctr = np.array([40.01, -74.99]) + rng.uniform(-0.0085, 0.0085, size=2)
        half = rng.uniform(0.0004, 0.0018)
        th = rng.uniform(0, 2 * np.pi)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        corners[i] = ctr + (np.array([[1, -1], [1, 1], [-1, 1], [-1, -1]]) * half) @ R.T
        dirs[i] = round(mo.get_direction(np.mean(corners[i], axis=0), (corners[i][0] + corners[i][1]) / 2)) % 360
        if i % 5 == 0:
            dirs[i] = (dirs[i] + 40) % 360                 # heading-correction branch
    bounds = np.tile(np.concatenate([bl, tr]), (B, 1))
    return corners, dirs, bounds


@pytest.mark.parametrize("last", [False, True])
def test_waypoint_step_kernel_vs_oracle(built_lib, last):
    from avdn_b200 import _lib
    B = 2048
    corners, dirs, bounds = _poses(B, 7)
    g = torch.Generator().manual_seed(1)
    out = torch.randn(B, 4, generator=g) * 1.2
    out[:, 3] = torch.rand(B, generator=g) * 0.6
    ended0 = (torch.rand(B, generator=g) < 0.2)
    ref_c, ref_d, ref_e, ref_a, ref_alt, ref_dist = mo.waypoint_step(out.numpy(), corners, bounds, dirs, ended0.numpy(), 0.25, last)
    d = "cuda"
    co = torch.from_numpy(corners).to(d).contiguous()
    cd = torch.from_numpy(dirs).to(d)
    en = ended0.to(torch.uint8).to(d)
    bd = torch.from_numpy(bounds).to(d).contiguous()
    o = out.to(d).contiguous()
    ang = torch.empty(B, dtype=torch.int32, device=d)
    alt = torch.empty(B, dtype=torch.int32, device=d)
    dist = torch.empty(B, dtype=torch.float64, device=d)
    ptr = _lib.ptr
    _lib.call("avdn_waypoint_step", ptr(o), ptr(co), ptr(bd), ptr(cd), ptr(en), B, 0.25, int(last), ptr(ang), ptr(dist),
              ptr(alt))
    assert np.array_equal(ang.cpu().numpy().astype(np.int64), ref_a)
    assert np.array_equal(alt.cpu().numpy().astype(np.int64), ref_alt)
    assert np.array_equal(en.cpu().numpy().astype(bool), ref_e)
    assert np.array_equal(cd.cpu().numpy(), ref_d)                         # integer headings, exact
    np.testing.assert_allclose(dist.cpu().numpy(), ref_dist, rtol=1e-12)
    np.testing.assert_allclose(co.cpu().numpy(), ref_c, rtol=1e-12, atol=0)
    moved = ~np.all(ref_c == corners, axis=(1, 2))
    assert (~moved).any() and (moved.any() != last)          # nobody moves on the last step


@pytest.fixture(scope="module")
def agent(built_lib):
    from avdn_b200.xview_lstm.agent import NavCMTAgent
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(mo.yolov3_trunk_cfg())
    torch.manual_seed(0)
    a = NavCMTAgent(types.SimpleNamespace(darknet_model_file=f.name, darknet_weight_file=None, max_action_len=4),
                    device="cuda")
    os.unlink(f.name)
    tile = wo.synthetic_tile(seed=2, size=1024)
    a.renderer.add_map("m", tile, None)
    a._tile = tile
    return a


def test_greedy_rollout_on_device_vs_oracle_pipeline(agent):
    """B=4, T=4 on a 1024x1024 tile.  Step 0 end to end against the oracle pipeline (cv2-exact views ->
    fp32 trunk in eval mode -> ViT_LSTM step); every step's simulator update against the oracle fed with
    OUR network outputs (teacher forcing: the discretisation may legitimately flip between an fp32 and a
    bf16 trunk, the update rule may not)."""
    B, T, L = 4, 4, 9
    size = 1024
    lat_ratio = 0.02 / size
    geo = np.tile(np.array([40.0, -75.0, 40.02, -74.98, lat_ratio]), (B, 1))
    corners, dirs, bounds = _poses(B, 3)
    g = torch.Generator().manual_seed(2)
    lang = torch.randn(B, L, 768, generator=g)
    cls = torch.relu(torch.randn(B, 49, generator=g))
    batch = dict(corners_gps=torch.from_numpy(corners).cuda(), directions=torch.from_numpy(dirs).cuda(),
                 geo=torch.from_numpy(geo).cuda(), tile_idx=None, lang_feature=lang.cuda(), cls_hidden=cls.cuda())
    res = agent.rollout_greedy(batch, max_action_len=T)
    torch.cuda.synchronize()
    assert agent.launches > 0
    out = res["output"].cpu()
    ch = res["corners"].cpu().numpy()
    dh = res["directions"].cpu().numpy()
    eh = res["ended"].cpu().numpy().astype(bool)
    assert np.array_equal(ch[0], corners) and np.array_equal(dh[0], dirs)
    # ---- step 0 through the oracle pipeline ----
    px = np.zeros((B, 4, 2), dtype=np.int32)
    for i in range(B):
        for k in range(4):
            lat, lng = corners[i, k]
            px[i, k] = (int(round((lng - geo[i, 1]) / lat_ratio)), int(round((geo[i, 2] - lat) / lat_ratio)))
    views = np.stack([wo.warp_fixed_point(agent._tile, wo.inverse_homography(px[i])) for i in range(B)])
    x = torch.from_numpy(wo.normalise_views(views))
    sd_t = {k: v.detach().cpu() for k, v in agent.vision_model.state_dict().items()}
    sd_l = {k: v.detach().cpu() for k, v in agent.vln_model.state_dict().items() if not k.startswith("vision_model.")}
    with torch.no_grad():
        feats = mo.darknet_forward(x, sd_t, mo.yolov3_trunk_cfg(), train=False).view(B, 512, 49)
        r = mo.vit_lstm_step(sd_l, feats, torch.from_numpy(dirs).long().view(B, 1), cls, lang)
    assert _rel(out[0], r[4]) < 2e-2, _rel(out[0], r[4])           # bf16 trunk (eval-mode BN: well conditioned)
    # ---- every step's simulator update, teacher-forced ----
    ended = np.zeros(B, dtype=bool)
    for t in range(T):
        nc, nd, ended, ang, alt, dist = mo.waypoint_step(out[t].numpy(), ch[t], bounds, dh[t], ended, 0.25, t == T - 1)
        assert np.array_equal(res["angle"][t].cpu().numpy().astype(np.int64), ang), t
        assert np.array_equal(res["altitude"][t].cpu().numpy().astype(np.int64), alt), t
        assert np.array_equal(eh[t], ended), t
        assert np.array_equal(dh[t + 1], nd), t
        np.testing.assert_allclose(ch[t + 1], nc, rtol=1e-12, atol=0)
    assert eh[T - 1].all()
    traj = agent.trajectories(res)
    assert len(traj) == B and all(len(p) >= 1 for p in traj)


@pytest.mark.parametrize("M,N,K,act,acc,pitch", [
    (256, 2304, 576, 0, 1, 0), (256, 768, 1536, 2, 0, 0), (256, 768, 768, 1, 0, 768), (256, 768, 192, 0, 1, 0),
    (256, 4, 32, 0, 0, 0), (3, 64, 256, 1, 0, 0), (37, 70, 100, 0, 0, 4), (1, 1, 4, 2, 1, 0), (65, 129, 36, 1, 1, 8)])
def test_linear_f32_v2_bit_identical_to_v1(built_lib, M, N, K, act, acc, pitch):
    """avdn_lstm_set_kernels: the register-prefetching 32 x 64-tile linear accumulates every output over ascending k
    with fmaf exactly as the first kernel does -> bit-identical results; both against torch float64 (1e-5).  Shapes:
    the rollout's own (256 episodes; K-offset views of wider buffers), ragged tiles in M and N, K not a multiple of
    the k-step, a single output."""
    from avdn_b200 import _lib
    call, ptr = _lib.call, _lib.ptr
    h = _lib.lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M * 131 + N * 7 + K)
    xbuf = torch.randn(M, K + pitch, device=dev, generator=g)
    x = xbuf[:, pitch:]                                   # a view whose rows start `pitch` floats into a wider buffer
    w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    y0 = torch.randn(M, N, device=dev, generator=g)

    def run(version):
        old = h.avdn_lstm_set_kernels(version)
        try:
            assert h.avdn_lstm_set_kernels(0) == version
            y = y0.clone()
            call("avdn_linear_f32", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(b), ptr(y), y.stride(0), M, N, K,
                 act, acc)
            torch.cuda.synchronize()
            return y
        finally:
            h.avdn_lstm_set_kernels(old)

    y1, y2 = run(1), run(2)
    assert torch.equal(y1, y2)
    ref = x.double() @ w.double().T + b.double() + (y0.double() if acc else 0)
    ref = torch.relu(ref) if act == 1 else (torch.tanh(ref) if act == 2 else ref)
    assert _rel(y2, ref) < 1e-5, _rel(y2, ref)


@pytest.mark.parametrize("B,L,D", [(256, 250, 768), (3, 7, 49), (5, 33, 1000), (2, 250, 1024), (4, 1, 256), (2, 17, 4),
                                   (3, 40, 1028)])
def test_lang_attn_v2_vs_v1_and_fp64(built_lib, B, L, D):
    """The second SoftDotAttention kernel (float4 loads, two / eight rows in flight; D % 4 == 0, D <= 1024 -- other
    shapes fall back to the first) sums each row in another lane order: against the first kernel and against float64
    to fp32 summation noise."""
    from avdn_b200 import _lib
    call, ptr = _lib.call, _lib.ptr
    h = _lib.lib()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(B + L + D)
    ctx = torch.randn(B, L, D, device=dev, generator=g)
    tgt = torch.randn(B, D, device=dev, generator=g) / D ** 0.5

    def run(version):
        old = h.avdn_lstm_set_kernels(version)
        try:
            attn = torch.zeros(B, L, device=dev)
            out = torch.zeros(B, 2 * D, device=dev)
            call("avdn_lang_attn_fwd", ptr(ctx), ptr(tgt), B, L, D, ptr(attn), ptr(out), out.stride(0))
            torch.cuda.synchronize()
            return attn, out[:, :D].clone()
        finally:
            h.avdn_lstm_set_kernels(old)

    (a1, o1), (a2, o2) = run(1), run(2)
    assert _rel(a2, a1) < 1e-5 and _rel(o2, o1) < 1e-5
    if D % 4:
        assert torch.equal(a1, a2) and torch.equal(o1, o2)          # fallback: the same kernel
    p = torch.softmax(torch.einsum("bld,bd->bl", ctx.double(), tgt.double()), dim=1)
    assert _rel(a2, p) < 1e-5
    assert _rel(o2, torch.einsum("bl,bld->bd", p, ctx.double())) < 1e-5
