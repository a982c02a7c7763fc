"""GPU: map preparation (src/env.py:217-231) on the device against cv2 itself --
INTER_AREA width rescale, filled attention circles, and the observation interface on a device-prepared map."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from oracle import map_oracle as mpo
from oracle import warp_oracle as wo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer(built_lib):
    from avdn_b200.env import ViewRenderer
    return ViewRenderer("cuda")


@pytest.mark.parametrize("H,W,lng,lat", [(300, 640, 0.7593e-5, 1e-5), (123, 1000, 2.2e-6, 3.3e-6), (64, 257, 1.0, 1.0),
                                         (200, 3000, 8.1e-6, 1.07e-5)])
def test_prepare_map_matches_cv2(renderer, H, W, lng, lat):
    rng = np.random.default_rng(H + W)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    new_w = int(W * lng / lat)
    spots = [((int(rng.integers(-20, new_w + 20)), int(rng.integers(-20, H + 20))), int(rng.integers(0, 90)))
             for _ in range(int(rng.integers(0, 8)))]
    idx, shape, m_host, a_host = renderer.prepare_map(f"m{H}x{W}", im, lng, lat, spots, keep_host_copies=True)
    ref = cv2.resize(im, (new_w, H), interpolation=cv2.INTER_AREA) if new_w != W else im
    att = np.zeros((H, new_w, 3), np.uint8)
    for c, r in spots:
        cv2.circle(att, center=c, radius=r, color=(255, 255, 255), thickness=-1)
    assert shape == (H, new_w, 3)
    assert np.array_equal(m_host, ref)
    assert np.array_equal(a_host, att)
    assert np.array_equal(a_host, mpo.attention_map(H, new_w, spots))
    renderer.remove_map(f"m{H}x{W}")


def test_observation_on_a_device_prepared_map(built_lib):
    """load_map + _get_obs == the reference pipeline (cv2 resize + circles, then the cv2-exact warp oracle)."""
    from avdn_b200.env import ANDHNavBatch
    H, W = 700, 900
    rng = np.random.default_rng(5)
    im = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    lat_ratio, lng_ratio = 1.0e-5, 0.8e-5
    new_w = int(W * lng_ratio / lat_ratio)
    bl = np.array([40.0, -75.0])
    tr = np.array([40.0 + H * lat_ratio, -75.0 + new_w * lat_ratio])
    def gps(x, y):
        return [tr[0] - y * lat_ratio, bl[1] + x * lat_ratio]
    corners = [gps(200, 150), gps(420, 190), gps(380, 410), gps(160, 370)]
    item = dict(map_name="mm", route_index="0", gps_botm_left=bl, gps_top_right=tr, lng_ratio=lng_ratio,
                lat_ratio=lat_ratio, angle=0, gt_path_corners=[corners], instructions="", pre_dialogs="",
                attention_list=[(gps(300, 280), 60), (gps(500, 100), 35)])
    env = ANDHNavBatch(batch_size=1, device="cuda")
    env.batch = [item]
    shape = env.load_map("mm", im, item)
    assert shape == (H, new_w, 3)
    ob = env._get_obs()[0]
    ref_map = cv2.resize(im, (new_w, H), interpolation=cv2.INTER_AREA)
    ref_att = np.zeros((H, new_w, 3), np.uint8)
    for a in item["attention_list"]:
        cv2.circle(ref_att, center=env.gps_to_img_coords(a[0], item), radius=a[1], color=(255, 255, 255), thickness=-1)
    cpx = np.array([env.gps_to_img_coords(c, item) for c in corners], dtype=np.int32)
    assert np.array_equal(ob["current_view"], wo.render_view(ref_map, cpx))
    assert np.array_equal(ob["gt_saliency"], wo.gt_saliency_from_view(wo.render_view(ref_att, cpx)))
    assert ob["map_size"] == (H, new_w, 3)
    env.drop_unused_maps([])
    assert not env.renderer.has_map("mm")
