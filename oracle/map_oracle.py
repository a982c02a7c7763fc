"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the map preparation step that precedes the renderer
(src/env.py:217-231): ``cv2.resize(im, (int(W*lng_ratio/lat_ratio), H), INTER_AREA)`` and the
human-attention raster ``cv2.circle(att, center, radius, (255,255,255), thickness=-1)``.

Both algorithms live in OpenCV (third party; the reference pins opencv-python==4.6.0.66, the build container has
4.13): they are restated from OpenCV's resize.cpp (``computeResizeAreaTab`` + ``ResizeArea_``: float weights, float
accumulation in table order, ``saturate_cast<uchar>`` = round half to even) and drawing.cpp (``Circle``: midpoint
algorithm filling horizontal spans), and pinned against ``cv2`` itself in tests/test_map_oracle.py.
Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import this module.
"""
import math

import numpy as np


def area_table(ssize, dsize):
    """OpenCV's computeResizeAreaTab for one axis: list of (dst index, src index, float32 weight)."""
    scale = ssize / dsize                       # double
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_width(im, new_w):
    """cv2.resize(im, (new_w, H), interpolation=cv2.INTER_AREA) for new_w < W (horizontal shrink, rows kept)."""
    H, W, C = im.shape
    assert new_w <= W
    if new_w == W:
        return im.copy()
    tab = area_table(W, new_w)
    acc = np.zeros((H, new_w, C), np.float32)
    src = im.astype(np.float32)
    for dx, sx, a in tab:                        # float32 accumulation in table order, as ResizeArea_ does
        acc[:, dx] = acc[:, dx] + src[:, sx] * a
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def circle_spans(radius):
    """Half-widths of the filled circle per row offset: hw[d] for d in [0, radius] (rows cy-d and cy+d), following
    OpenCV's ``Circle`` (drawing.cpp): every iteration fills rows cy +- dy with half-width dx and rows cy +- dx with
    half-width dy."""
    hw = np.full(radius + 1, -1, np.int64)
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1
        err += plus
        plus += 2
        mask = (1 if err <= 0 else 0) - 1        # 0 or -1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return hw


def filled_circle(img, center, radius, value=255):
    """cv2.circle(img, center, radius, (value,)*3, thickness=-1) in place (any channel count)."""
    H, W = img.shape[:2]
    cx, cy = int(center[0]), int(center[1])
    hw = circle_spans(int(radius))
    for d in range(int(radius) + 1):
        if hw[d] < 0:
            continue
        for y in (cy - d, cy + d):
            if 0 <= y < H:
                x0, x1 = max(cx - hw[d], 0), min(cx + hw[d], W - 1)
                if x0 <= x1:
                    img[y, x0:x1 + 1] = value
    return img


def attention_map(H, W, spots):
    """src/env.py:224-230: zeros [H,W,3] u8 + one filled white circle per (center (x,y), radius)."""
    att = np.zeros((H, W, 3), np.uint8)
    for (c, r) in spots:
        filled_circle(att, c, r, 255)
    return att
