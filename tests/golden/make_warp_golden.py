"""Generates tests/golden/warp_golden.npz by running the REFERENCE's own
``ANDHNavBatch._get_obs`` (src/env.py:254-332, i.e. cv2.getPerspectiveTransform +
cv2.warpPerspective) in the build container.  /root/reference is read-only and
absent on the GPU box, hence the committed fixture.

    python tests/golden/make_warp_golden.py

Shims: shapely is not installed, env.py imports it at module
scope but the hot path never calls it -> stub modules.
"""
import hashlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import warp_oracle as wo  # noqa: E402

for name in ("shapely", "shapely.geometry", "shapely.ops"):
    m = types.ModuleType(name)
    for sym in ("Point", "Polygon", "LineString", "MultiPoint", "Polygon", "nearest_points"):
        setattr(m, sym, object)
    sys.modules[name] = m
sys.path.insert(0, "/root/reference/src")
import cv2  # noqa: E402
import env as ref_env  # noqa: E402

SIZE = 640
N = 12


def main():
    tile = wo.synthetic_tile(seed=7, size=SIZE)
    att = wo.synthetic_attention_tile(seed=7, size=SIZE)
    # small discs so that some views see attention and some do not
    att[:] = 0
    yy, xx = np.mgrid[0:SIZE, 0:SIZE]
    for cx, cy, r in ((200, 220, 60), (450, 380, 40), (90, 560, 25)):
        att[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = 255
    rng = np.random.default_rng(11)
    # pixel-space footprints -> GPS, so that gps_to_img_coords is exercised too
    bl = np.array([34.0, -118.0])                 # (lat, lng) of the bottom-left
    lat_ratio = 2.7e-6
    tr = np.array([bl[0] + SIZE * lat_ratio, bl[1] + SIZE * lat_ratio])
    px = wo.synthetic_pose_corners(N, seed=5, size=3000, edge_frac=0.0).astype(np.float64)
    px = (px - 1500) * 0.2 + SIZE / 2            # footprints 27..267 px inside a 640 tile
    px[-3:] += rng.uniform(200, 330, size=(3, 1, 2))   # three poses hang off the tile
    px += rng.uniform(-0.5, 0.5, size=px.shape)  # non-integer -> rounding matters
    gps = np.stack([tr[0] - px[..., 1] * lat_ratio, bl[1] + px[..., 0] * lat_ratio], -1)

    e = ref_env.ANDHNavBatch.__new__(ref_env.ANDHNavBatch)
    e.batch_size = N
    e.map_batch = {"m0": tile}
    e.attention_map_batch = {"m0": att}
    e.batch = [dict(map_name="m0", route_index=str(i), gps_botm_left=bl, gps_top_right=tr,
                    lng_ratio=lat_ratio * 1.2, lat_ratio=lat_ratio, angle=0,
                    gt_path_corners=[gps[i].copy()], instructions="", pre_dialogs=[])
               for i in range(N)]
    obs = e._get_obs(t=0)
    views = np.stack([o["current_view"] for o in obs])
    sal = np.stack([o["gt_saliency"] for o in obs])
    corners_px = np.stack([o["view_area_corners"] for o in obs])
    dst = np.array([[0, 0], [223, 0], [223, 223], [0, 223]], dtype=np.float32)
    M = np.stack([cv2.getPerspectiveTransform(c.astype(np.float32), dst) for c in corners_px])
    Mi = np.stack([cv2.invert(m)[1] for m in M])
    out = dict(
        size=SIZE, tile_seed=7,
        tile_sha=hashlib.sha256(tile.tobytes()).hexdigest(),
        att_tile=np.packbits(att[:, :, 0] > 0),
        gps=gps, bl=bl, tr=tr, lat_ratio=lat_ratio,
        corners_px=corners_px.astype(np.int32), M=M, Minv=Mi,
        views_sha=np.array([hashlib.sha256(v.tobytes()).hexdigest() for v in views]),
        views_sub=views[:, ::7, ::7].copy(),
        sal_u8=np.rint(sal * 255).astype(np.uint8)[:, ::3, ::3].copy(),
        sal_sha=np.array([hashlib.sha256(np.rint(s * 255).astype(np.uint8).tobytes()).hexdigest() for s in sal]),
        sal_sum=sal.sum(axis=(1, 2)),
        cv2_version=cv2.__version__,
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "warp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; sal sums", out["sal_sum"])


if __name__ == "__main__":
    main()
