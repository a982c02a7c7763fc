"""CPU: the renderer oracle against the reference-generated golden vectors and,
when cv2 is importable, against cv2 live (the library the reference calls at
src/env.py:287,290,292)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import warp_oracle as wo


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "warp_golden.npz"))


def _golden_tiles(g):
    size = int(g["size"])
    tile = wo.synthetic_tile(seed=int(g["tile_seed"]), size=size)
    assert hashlib.sha256(tile.tobytes()).hexdigest() == str(g["tile_sha"])
    att = (np.unpackbits(g["att_tile"])[: size * size].reshape(size, size) * 255).astype(np.uint8)
    return tile, att


def test_gps_to_pixels_matches_reference(golden):
    px = wo.gps_corners_to_pixels(golden["gps"], golden["bl"][None], golden["tr"][None],
                                  np.array([float(golden["lat_ratio"])]))
    assert np.array_equal(px, golden["corners_px"])
    # scalar version, incl. half-to-even
    for p in range(px.shape[0]):
        for k in range(4):
            assert wo.gps_to_img_coords(golden["gps"][p, k], golden["bl"], golden["tr"],
                                        float(golden["lat_ratio"])) == tuple(px[p, k])
    assert wo.gps_to_img_coords((0.0, 2.5), (0, 0), (0, 0), 1.0)[0] == 2
    assert wo.gps_to_img_coords((0.0, 3.5), (0, 0), (0, 0), 1.0)[0] == 4


def test_homography_bit_exact(golden):
    for c, M, Mi in zip(golden["corners_px"], golden["M"], golden["Minv"]):
        Mo = wo.perspective_transform(c.astype(np.float32))
        assert np.array_equal(Mo, M)
        assert np.array_equal(wo.invert3x3(Mo), Mi)


def test_views_and_saliency_bit_exact(golden):
    tile, att = _golden_tiles(golden)
    for i, c in enumerate(golden["corners_px"]):
        Mi = wo.inverse_homography(c)
        v = wo.warp_fixed_point(tile, Mi)
        assert hashlib.sha256(v.tobytes()).hexdigest() == str(golden["views_sha"][i])
        assert np.array_equal(v[::7, ::7], golden["views_sub"][i])
        a = wo.warp_fixed_point(att, Mi)[:, :, 0]
        assert hashlib.sha256(a.tobytes()).hexdigest() == str(golden["sal_sha"][i])
        assert np.isclose(wo.gt_saliency_from_view(a).sum(), golden["sal_sum"][i], rtol=0, atol=1e-9)


def test_against_cv2_live():
    cv2 = pytest.importorskip("cv2")
    tile = wo.synthetic_tile(seed=3, size=900)
    corners = wo.synthetic_pose_corners(24, seed=2, size=3000, edge_frac=0.0).astype(np.float64)
    corners = np.rint((corners - 1500) * 0.3 + 450).astype(np.int32)
    corners[-4:] += 500                       # off-tile: BORDER_CONSTANT path
    dst = np.array([[0, 0], [223, 0], [223, 223], [0, 223]], dtype=np.float32)
    for c in corners:
        M = cv2.getPerspectiveTransform(c.astype(np.float32), dst)
        ref = cv2.warpPerspective(tile, M, (224, 224))
        assert np.array_equal(wo.perspective_transform(c.astype(np.float32)), M)
        assert np.array_equal(wo.render_view(tile, c), ref)


def test_degenerate_quad_is_pinned():
    """Zero-area footprints: pinned to the all-zero inverse (see warp_oracle.inverse_homography)."""
    tile = wo.synthetic_tile(seed=4, size=300)
    c = np.array([[10, 10], [50, 50], [90, 90], [130, 130]], dtype=np.int32)   # collinear
    assert np.array_equal(wo.inverse_homography(c), np.zeros((3, 3)))
    v = wo.render_view(tile, c)
    assert (v == tile[0, 0]).all()


def test_normalise_matches_reference_snippet():
    rng = np.random.default_rng(0)
    v = rng.integers(0, 256, size=(2, 224, 224, 3), dtype=np.uint8)
    out = wo.normalise_views(v)
    assert out.shape == (2, 3, 224, 224) and out.dtype == np.float32
    r = (v[0, 5, 7, 2].astype(np.float32) - np.float32(60.134)) / np.float32(29.99)
    assert out[0, 0, 5, 7] == r
