"""CPU oracle for the AVDN view renderer (stage 1 of the hot path).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as
the checker / the CPU baseline.

What it restates (reference ``file:line`` relative to /root/reference):

* ``gps_to_img_coords``      src/env.py:189-196  (Python ``round`` = half-to-even)
* corner prep + homography   src/env.py:273-287  (``cv2.getPerspectiveTransform``)
* the two warps              src/env.py:290,292  (``cv2.warpPerspective``,
                             INTER_LINEAR, BORDER_CONSTANT 0, uint8)
* gray / 255                 src/env.py:293
* image normalisation        src/xview_et/agent.py:586-592 (consts :115-116)

The arithmetic of the warp itself lives in a third-party dependency that is
not vendored in /root/reference: OpenCV (the reference pins
``opencv-python==4.6.0.66`` in requirements.txt:14; this container has 4.13.0).
The published algorithm restated here is OpenCV's fixed-point remap path
(``imgwarp.cpp``: ``getPerspectiveTransform`` -> 8x8 LU with partial pivoting in
float64, ``invert`` 3x3 closed form, ``WarpPerspectiveInvoker`` 64-wide blocks,
``INTER_BITS=5``, ``INTER_REMAP_COEF_BITS=15``).

Parity pin: the reference ships no golden vectors for this path (SURVEY.md §4,
§8c).  The oracle is pinned instead against the reference's own call
(``cv2.getPerspectiveTransform`` + ``cv2.warpPerspective``) executed in the build
container: ``tests/golden/make_warp_golden.py`` generated the committed
fixtures, and ``tests/test_warp_oracle.py`` re-checks the oracle against cv2
live whenever cv2 is importable.
"""
from __future__ import annotations

import numpy as np

VIEW = 224                      # src/env.py:273-274
INTER_BITS = 5
INTER_TAB = 1 << INTER_BITS     # 32
BLOCK_W = 64                    # OpenCV block width for a 224-wide destination

RGB_MEAN = np.array([60.134, 49.697, 40.746], dtype=np.float32)   # agent.py:115
RGB_STD = np.array([29.99, 24.498, 22.046], dtype=np.float32)     # agent.py:116


def gps_to_img_coords(gps, gps_botm_left, gps_top_right, lat_ratio):
    """src/env.py:189-196.  Both axes divide by ``lat_ratio`` (reference quirk).

    ``gps`` = (lat, lng).  Returns (x, y) ints; rounding is half-to-even on
    float64, as Python's ``round``.
    """
    x = int(round((gps[1] - gps_botm_left[1]) / lat_ratio))
    y = int(round((gps_top_right[0] - gps[0]) / lat_ratio))
    return x, y


def gps_corners_to_pixels(corners_gps, gps_botm_left, gps_top_right, lat_ratio):
    """Vectorised ``gps_to_img_coords`` over ``[P,4,2]`` (lat,lng) float64 corners.

    ``np.rint`` is round-half-to-even, identical to Python ``round`` for values
    that fit an int.  Returns int32 ``[P,4,2]`` (x, y).
    """
    c = np.asarray(corners_gps, dtype=np.float64)
    bl = np.asarray(gps_botm_left, dtype=np.float64).reshape(-1, 1, 2)
    tr = np.asarray(gps_top_right, dtype=np.float64).reshape(-1, 1, 2)
    lr = np.asarray(lat_ratio, dtype=np.float64).reshape(-1, 1)
    x = np.rint((c[..., 1] - bl[..., 1]) / lr)
    y = np.rint((tr[..., 0] - c[..., 0]) / lr)
    return np.stack([x, y], axis=-1).astype(np.int32)


def perspective_transform(src4, dst4=None):
    """Restatement of ``cv2.getPerspectiveTransform`` (call site src/env.py:287).

    8x8 system in float64, Gaussian elimination with partial pivoting exactly in
    OpenCV's ``LUImpl`` order (``d = -1/pivot``; ``row_j += (a_ji*d)*row_i``;
    back-substitution ``s -= a_ik*x_k`` then ``s / a_ii``).  No FMA contraction
    (plain Python floats are IEEE double, one rounding per operation).
    """
    src = np.asarray(src4, dtype=np.float32).astype(np.float64)
    if dst4 is None:
        w = h = VIEW
        dst4 = [[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]]    # env.py:275-278
    dst = np.asarray(dst4, dtype=np.float32).astype(np.float64)
    a = [[0.0] * 8 for _ in range(8)]
    b = [0.0] * 8
    for i in range(4):
        sx, sy = float(src[i, 0]), float(src[i, 1])
        dx, dy = float(dst[i, 0]), float(dst[i, 1])
        a[i][0] = a[i + 4][3] = sx
        a[i][1] = a[i + 4][4] = sy
        a[i][2] = a[i + 4][5] = 1.0
        a[i][6] = -sx * dx
        a[i][7] = -sy * dx
        a[i + 4][6] = -sx * dy
        a[i + 4][7] = -sy * dy
        b[i] = dx
        b[i + 4] = dy
    m = 8
    eps = np.finfo(np.float64).eps * 100
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(a[j][i]) > abs(a[k][i]):
                k = j
        if abs(a[k][i]) < eps:
            return None            # singular: OpenCV returns a zero matrix solution
        if k != i:
            a[i], a[k] = a[k], a[i]
            b[i], b[k] = b[k], b[i]
        d = -1.0 / a[i][i]
        for j in range(i + 1, m):
            alpha = a[j][i] * d
            for kk in range(i + 1, m):
                a[j][kk] += alpha * a[i][kk]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for kk in range(i + 1, m):
            s -= a[i][kk] * b[kk]
        b[i] = s / a[i][i]
    return np.array(b + [1.0], dtype=np.float64).reshape(3, 3)


def invert3x3(M):
    """Restatement of ``cv::invert`` (DECOMP_LU) for a 3x3 float64 matrix: the
    closed-form adjugate / determinant path OpenCV takes for n == 3."""
    s = [[float(M[r, c]) for c in range(3)] for r in range(3)]
    det = (s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1])
           - s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0])
           + s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]))
    if det == 0.0:
        return None
    d = 1.0 / det
    t = [0.0] * 9
    t[0] = (s[1][1] * s[2][2] - s[1][2] * s[2][1]) * d
    t[1] = (s[0][2] * s[2][1] - s[0][1] * s[2][2]) * d
    t[2] = (s[0][1] * s[1][2] - s[0][2] * s[1][1]) * d
    t[3] = (s[1][2] * s[2][0] - s[1][0] * s[2][2]) * d
    t[4] = (s[0][0] * s[2][2] - s[0][2] * s[2][0]) * d
    t[5] = (s[0][2] * s[1][0] - s[0][0] * s[1][2]) * d
    t[6] = (s[1][0] * s[2][1] - s[1][1] * s[2][0]) * d
    t[7] = (s[0][1] * s[2][0] - s[0][0] * s[2][1]) * d
    t[8] = (s[0][0] * s[1][1] - s[0][1] * s[1][0]) * d
    return np.array(t, dtype=np.float64).reshape(3, 3)


def inverse_homography(corners_px):
    """int pixel corners (FL,FR,BR,BL) ``[4,2]`` -> inverse homography ``Mi`` (3x3
    float64) that ``warpPerspective`` evaluates per destination pixel."""
    M = perspective_transform(np.asarray(corners_px, dtype=np.float32))
    Mi = invert3x3(M) if M is not None else None
    if Mi is None:
        # Degenerate (zero-area) footprint.  OpenCV 4.13 switches to another
        # solver here and returns an arbitrary near-singular matrix; such poses
        # are meaningless for the simulator, so both this oracle and the CUDA
        # kernel pin the behaviour to "all-zero inverse" (every pixel then
        # samples tile[0][0]).  Documented deviation, DESIGN.md.
        return np.zeros((3, 3), dtype=np.float64)
    return Mi


def warp_fixed_point(tile, Mi, width=VIEW, height=VIEW):
    """Restatement of ``cv2.warpPerspective(tile, M, (w,h))`` with
    ``Mi = invert(M)`` for uint8 HWC tiles (src/env.py:290,292).

    Integer result; the tap weights are products of two 5-bit fractions, the
    output is ``(sum + 512) >> 10``, taps outside the tile read 0.
    """
    tile = np.asarray(tile)
    if tile.ndim == 2:
        tile = tile[:, :, None]
    H, W, C = tile.shape
    m = np.asarray(Mi, dtype=np.float64)
    x = np.arange(width, dtype=np.int64)
    xb = (x // BLOCK_W) * BLOCK_W
    x1 = (x - xb).astype(np.float64)
    xbf = xb.astype(np.float64)
    y = np.arange(height, dtype=np.float64)[:, None]
    # float64, evaluation order of WarpPerspectiveInvoker: (M0*xb + M1*y) + M2
    X0 = (m[0, 0] * xbf[None, :] + m[0, 1] * y) + m[0, 2]
    Y0 = (m[1, 0] * xbf[None, :] + m[1, 1] * y) + m[1, 2]
    W0 = (m[2, 0] * xbf[None, :] + m[2, 1] * y) + m[2, 2]
    Wd = W0 + m[2, 0] * x1[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        Wd = np.where(Wd != 0.0, INTER_TAB / Wd, 0.0)
        fX = (X0 + m[0, 0] * x1[None, :]) * Wd
        fY = (Y0 + m[1, 0] * x1[None, :]) * Wd
    imin, imax = float(np.iinfo(np.int32).min), float(np.iinfo(np.int32).max)
    fX = np.maximum(imin, np.minimum(imax, fX))
    fY = np.maximum(imin, np.minimum(imax, fY))
    X = np.rint(fX).astype(np.int64)          # round-half-even, as cvRound (SSE2)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)   # saturate_cast<short>
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    ax = (X & (INTER_TAB - 1)).astype(np.int64)
    ay = (Y & (INTER_TAB - 1)).astype(np.int64)

    def tap(r, c):
        ok = (r >= 0) & (r < H) & (c >= 0) & (c < W)
        rr = np.clip(r, 0, H - 1)
        cc = np.clip(c, 0, W - 1)
        v = tile[rr, cc].astype(np.int64)
        return v * ok[..., None]

    w00 = ((INTER_TAB - ay) * (INTER_TAB - ax))[..., None]
    w01 = ((INTER_TAB - ay) * ax)[..., None]
    w10 = (ay * (INTER_TAB - ax))[..., None]
    w11 = (ay * ax)[..., None]
    acc = (tap(sy, sx) * w00 + tap(sy, sx + 1) * w01
           + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11)
    out = ((acc + 512) >> 10).astype(np.uint8)
    return out


def render_view(tile, corners_px):
    """One ``current_view`` (uint8 ``[224,224,C]``) from int pixel corners."""
    return warp_fixed_point(tile, inverse_homography(corners_px))


def gt_saliency_from_view(att_view):
    """src/env.py:293.  ``BGR2GRAY`` of an R=G=B image is the channel itself;
    the result is float64 in [0,1]."""
    a = np.asarray(att_view)
    if a.ndim == 3:
        a = a[:, :, 0]
    return a.astype(np.float64) / 255


def normalise_views(views_bgr_u8):
    """src/xview_et/agent.py:586-592: ``[B,224,224,3]`` BGR u8 -> ``[B,3,224,224]``
    RGB float32, ``(x - mean) / std`` as two separate float32 operations."""
    images = np.asarray(views_bgr_u8)[:, :, :, ::-1].transpose(0, 3, 1, 2)
    images = np.ascontiguousarray(images, dtype=np.float32)
    images -= RGB_MEAN.reshape(3, 1, 1)
    images /= RGB_STD.reshape(3, 1, 1)
    return images


# --------------------------------------------------------------------------
# synthetic inputs of the ANDH shape
# --------------------------------------------------------------------------
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
