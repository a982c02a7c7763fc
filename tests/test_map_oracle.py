"""CPU: the map-preparation restatement (oracle/map_oracle.py) pinned against cv2 itself -- the calls at
src/env.py:221 (INTER_AREA width rescale) and src/env.py:226-230 (filled attention circles)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import map_oracle as mpo


@pytest.mark.parametrize("W,ratio", [(640, 0.7593), (1000, 0.6631), (777, 0.9), (512, 0.5), (300, 0.3337), (257, 1.0)])
def test_area_resize_matches_cv2(W, ratio):
    rng = np.random.default_rng(W)
    im = rng.integers(0, 256, size=(37, W, 3), dtype=np.uint8)
    new_w = int(W * ratio)
    ref = cv2.resize(im, (new_w, im.shape[0]), interpolation=cv2.INTER_AREA)
    assert np.array_equal(mpo.resize_area_width(im, new_w), ref)


def test_filled_circles_match_cv2():
    rng = np.random.default_rng(0)
    for r in list(range(0, 40)) + [57, 100, 150, 211]:
        H, W = 260, 300
        cx, cy = int(rng.integers(-50, W + 50)), int(rng.integers(-50, H + 50))
        ref = np.zeros((H, W, 3), np.uint8)
        cv2.circle(ref, center=(cx, cy), radius=r, color=(255, 255, 255), thickness=-1)
        ours = mpo.attention_map(H, W, [((cx, cy), r)])
        assert np.array_equal(ours, ref), (r, cx, cy)
