"""Synthetic workload inputs of the ANDH shape: the satellite tile, its
human-attention map, rotated-square pose footprints, and the text of the (inferred) truncated
xview-yolov3 trunk cfg.  Used by ``bench.py`` and the examples; deterministic in ``seed``.

The test oracle (``oracle/``) carries its own copies of these generators so that it stays free of
product imports; ``tests/test_synthetic.py`` pins the two to identical outputs.
"""
import numpy as np


def synthetic_tile(seed=0, size=3000, smooth=False):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, size=(size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    t = np.zeros((size, size, 3), np.float32)
    for c in range(3):
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.002, 0.03), rng.uniform(0.002, 0.03), rng.uniform(0, 6.28)
            t[:, :, c] += np.sin(xx * fx + yy * fy + ph)
    t = (t - t.min()) / (t.max() - t.min()) * 255
    return t.astype(np.uint8)


def synthetic_attention_tile(seed=0, size=3000):
    """zeros + filled discs of value 255 on all 3 channels (src/env.py:224-230)."""
    rng = np.random.default_rng(seed + 1000)
    t = np.zeros((size, size, 3), np.uint8)
    yy, xx = np.mgrid[0:size, 0:size]
    for _ in range(int(rng.integers(3, 11))):
        cx, cy, r = rng.uniform(0, size), rng.uniform(0, size), rng.uniform(30, 150)
        t[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = 255
    return t


def synthetic_pose_corners(n, seed=0, size=3000, edge_frac=0.05):
    """``n`` rotated square footprints as int32 pixel corners ``[n,4,2]`` in the
    reference order (front-left, front-right, back-right, back-left).

    centre ~ U[400,2600]^2, side ~ U[133,1333] px (40-400 m at 0.3 m/px),
    heading ~ U{0..359} deg; ``edge_frac`` of the poses are shifted so that the
    footprint overlaps the tile edge (BORDER_CONSTANT path).
    """
    rng = np.random.default_rng(seed)
    c = rng.uniform(400, size - 400, size=(n, 2))
    side = rng.uniform(133, 1333, size=n)
    th = np.deg2rad(rng.integers(0, 360, size=n).astype(np.float64))
    edge = rng.random(n) < edge_frac
    c[edge] = rng.uniform(-100, size + 100, size=(int(edge.sum()), 2))
    fwd = np.stack([np.sin(th), -np.cos(th)], 1)          # heading 0 = up (row 0)
    right = np.stack([np.cos(th), np.sin(th)], 1)
    h = (side / 2)[:, None]
    fl = c + fwd * h - right * h
    fr = c + fwd * h + right * h
    br = c - fwd * h + right * h
    bl = c - fwd * h - right * h
    return np.rint(np.stack([fl, fr, br, bl], 1)).astype(np.int32)


def yolov3_trunk_cfg(truncate_after=None):
    """Text of the (inferred) truncated xview-yolov3 cfg: Darknet-53 + 5 head
    convs, ending at the 512-channel stride-32 layer."""
    out = ["[net]", "channels=3", "height=416", ""]

    def conv(f, k, s):
        out.extend(["[convolutional]", "batch_normalize=1", f"filters={f}", f"size={k}", f"stride={s}",
                    "pad=1", "activation=leaky", ""])

    def res(f, n):
        for _ in range(n):
            conv(f // 2, 1, 1)
            conv(f, 3, 1)
            out.extend(["[shortcut]", "from=-3", "activation=linear", ""])

    conv(32, 3, 1)
    conv(64, 3, 2); res(64, 1)
    conv(128, 3, 2); res(128, 2)
    conv(256, 3, 2); res(256, 8)
    conv(512, 3, 2); res(512, 8)
    conv(1024, 3, 2); res(1024, 4)
    conv(512, 1, 1); conv(1024, 3, 1); conv(512, 1, 1); conv(1024, 3, 1); conv(512, 1, 1)
    return "\n".join(out)
