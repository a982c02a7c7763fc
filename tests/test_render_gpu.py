"""GPU: the CUDA renderer (through the C ABI) against the oracle and the
reference-generated golden vectors.  Bit-exact everywhere."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(built_lib):
    from avdn_b200.env import ViewRenderer
    return ViewRenderer("cuda:0")


def test_golden_vectors(renderer, golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    size = int(g["size"])
    tile = wo.synthetic_tile(seed=int(g["tile_seed"]), size=size)
    att = (np.unpackbits(g["att_tile"])[: size * size].reshape(size, size) * 255).astype(np.uint8)
    att3 = np.repeat(att[:, :, None], 3, 2)
    renderer.add_map("golden", tile, att3)
    P = g["gps"].shape[0]
    geo = np.tile(np.concatenate([g["bl"], g["tr"], [float(g["lat_ratio"])]])[None], (P, 1))
    px = renderer.gps_to_pixels(g["gps"], geo)
    assert np.array_equal(px.cpu().numpy(), g["corners_px"])
    minv = renderer.homography(px)
    assert np.array_equal(minv.cpu().numpy(), g["Minv"])
    idx = torch.full((P,), renderer.map_index("golden"), dtype=torch.int32)
    out = renderer.render(px, idx, views=True, att=True)
    v, a = out["views"].cpu().numpy(), out["att"].cpu().numpy()
    for i in range(P):
        assert hashlib.sha256(v[i].tobytes()).hexdigest() == str(g["views_sha"][i])
        assert hashlib.sha256(a[i].tobytes()).hexdigest() == str(g["sal_sha"][i])
        assert np.isclose((a[i].astype(np.float64) / 255).sum(), g["sal_sum"][i], rtol=0, atol=1e-9)


def test_random_poses_vs_oracle(renderer):
    tile = wo.synthetic_tile(seed=21, size=1000)
    att = wo.synthetic_attention_tile(seed=21, size=1000)
    renderer.add_map("t1000", tile, att)
    c = wo.synthetic_pose_corners(48, seed=9, size=3000, edge_frac=0.0).astype(np.float64)
    c = np.rint((c - 1500) * 0.33 + 500).astype(np.int32)
    c[40:] += np.array([600, -350])          # off-tile (BORDER_CONSTANT)
    idx = torch.full((48,), renderer.map_index("t1000"), dtype=torch.int32)
    out = renderer.render(torch.from_numpy(c), idx, views=True, att=True, norm_nchw=True, norm_nhwc=True)
    v, a = out["views"].cpu().numpy(), out["att"].cpu().numpy()
    for i in range(48):
        Mi = wo.inverse_homography(c[i])
        assert np.array_equal(v[i], wo.warp_fixed_point(tile, Mi)), i
        assert np.array_equal(a[i], wo.warp_fixed_point(att, Mi)[:, :, 0]), i
    ref = wo.normalise_views(v)
    assert np.array_equal(out["norm_nchw"].cpu().numpy(), ref)
    nh = out["norm_nhwc"].float().cpu().numpy()
    exp = torch.from_numpy(ref).permute(0, 2, 3, 1).to(torch.bfloat16).float().numpy()
    assert np.array_equal(nh[..., :3], exp) and (nh[..., 3] == 0).all()


def test_axis_aligned_and_tie_heavy_poses(renderer):
    """Axis-aligned footprints whose scale is a multiple of 1/32 make many source
    coordinates land exactly on rounding ties: the exact fall-back path."""
    tile = wo.synthetic_tile(seed=5, size=700)
    renderer.add_map("t700", tile, None)
    cs = []
    for side, x0, y0 in ((223, 10, 20), (446, 100, 50), (112, 300, 300), (335, 5, 5), (669, 0, 0)):
        cs.append([[x0, y0], [x0 + side, y0], [x0 + side, y0 + side], [x0, y0 + side]])
        cs.append([[x0 + side, y0], [x0 + side, y0 + side], [x0, y0 + side], [x0, y0]])  # rotated 90
    c = np.array(cs, dtype=np.int32)
    idx = torch.full((len(c),), renderer.map_index("t700"), dtype=torch.int32)
    v = renderer.render(torch.from_numpy(c), idx)["views"].cpu().numpy()
    for i in range(len(c)):
        assert np.array_equal(v[i], wo.render_view(tile, c[i])), i


def test_degenerate_and_empty(renderer):
    tile = wo.synthetic_tile(seed=5, size=700)
    if not renderer.has_map("t700"):
        renderer.add_map("t700", tile, None)
    c = np.array([[[10, 10], [50, 50], [90, 90], [130, 130]]], dtype=np.int32)
    idx = torch.full((1,), renderer.map_index("t700"), dtype=torch.int32)
    v = renderer.render(torch.from_numpy(c), idx)["views"].cpu().numpy()
    assert (v[0] == tile[0, 0]).all()
    out = renderer.render(torch.zeros((0, 4, 2), dtype=torch.int32), torch.zeros((0,), dtype=torch.int32))
    assert out["views"].shape == (0, 224, 224, 3)


def test_full_size_properties(renderer):
    """Config 3 size (3000x3000 tile, 4096 poses): size-independent properties.
    (a) the identity-scale crop equals the tile slice; (b) a 180-degree rotated
    footprint renders the flipped view; (c) sampled poses equal the oracle."""
    tile = wo.synthetic_tile(seed=0, size=3000)
    renderer.add_map("full", tile, None)
    P = 4096
    c = wo.synthetic_pose_corners(P, seed=0, size=3000)
    c[0] = [[100, 200], [323, 200], [323, 423], [100, 423]]             # identity scale
    c[1] = c[0][[2, 3, 0, 1]]                                            # rotated by 180 deg
    idx = torch.full((P,), renderer.map_index("full"), dtype=torch.int32)
    v = renderer.render(torch.from_numpy(c), idx)["views"]
    assert np.array_equal(v[0].cpu().numpy(), tile[200:424, 100:324])
    assert np.array_equal(v[1].cpu().numpy(), tile[200:424, 100:324][::-1, ::-1])
    for i in (2, 77, 1234, 4095):
        assert np.array_equal(v[i].cpu().numpy(), wo.render_view(tile, c[i])), i
    renderer.remove_map("full")


def test_env_get_obs_interface(built_lib):
    """ANDHNavBatch._get_obs returns the reference's dict (src/env.py:296-319)."""
    from avdn_b200.env import ANDHNavBatch
    size = 640
    tile = wo.synthetic_tile(seed=7, size=size)
    att = np.repeat(wo.synthetic_attention_tile(seed=2, size=size)[:, :, :1], 3, 2)
    bl, lat_ratio = np.array([34.0, -118.0]), 2.7e-6
    tr = np.array([bl[0] + size * lat_ratio, bl[1] + size * lat_ratio])
    px = np.array([[[100.2, 120.7], [300.4, 110.1], [310.3, 320.9], [95.5, 330.5]],
                   [[400.0, 100.0], [500.0, 200.0], [400.0, 300.0], [300.0, 200.0]]])
    gps = np.stack([tr[0] - px[..., 1] * lat_ratio, bl[1] + px[..., 0] * lat_ratio], -1)
    env = ANDHNavBatch(batch_size=2)
    env.map_batch = {"m0": tile}
    env.attention_map_batch = {"m0": att}
    env.batch = [dict(map_name="m0", route_index=str(i), gps_botm_left=bl, gps_top_right=tr,
                      lng_ratio=lat_ratio, lat_ratio=lat_ratio, angle=0, gt_path_corners=[gps[i]],
                      instructions="go", pre_dialogs=[]) for i in range(2)]
    obs = env._get_obs(t=0)
    assert len(obs) == 2
    for i, o in enumerate(obs):
        cpx = wo.gps_corners_to_pixels(gps[i][None], bl[None], tr[None], np.array([lat_ratio]))[0]
        assert np.array_equal(o["view_area_corners"], cpx.astype(np.float64))
        assert o["current_view"].dtype == np.uint8 and o["current_view"].shape == (224, 224, 3)
        assert np.array_equal(o["current_view"], wo.render_view(tile, cpx))
        assert o["gt_saliency"].dtype == np.float64
        assert np.array_equal(o["gt_saliency"], wo.gt_saliency_from_view(wo.render_view(att, cpx)))
        assert o["map_size"] == tile.shape
    obs2 = env._get_obs(corners=[gps[1], gps[0]])
    assert np.array_equal(obs2[0]["current_view"], obs[1]["current_view"])
