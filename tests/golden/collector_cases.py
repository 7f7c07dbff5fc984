"""Inputs of the collector golden cases (shared by make_collector.py and the tests; needs no reference)."""
from __future__ import annotations

import cv2
import numpy as np

CONFIG_SEED = 7
RNG_SEED = 20240607
SMALL_CASES = 4
CRAFTED_CASES = 3


def crafted(idx: int):
    """360x480 mask + depth: an ellipse low in the image with one-pixel spikes."""
    H, W = 360, 480
    rng = np.random.default_rng(500 + idx)
    m = np.zeros((H, W), np.uint8)
    cx, cy = int(rng.integers(160, 320)), int(rng.integers(230, 270))
    cv2.ellipse(m, (cx, cy), (int(rng.integers(90, 120)), int(rng.integers(55, 75))), float(rng.uniform(0, 180)),
                0, 360, 1, -1)
    ys, xs = np.nonzero(m)
    top = ys.min()
    xt = int(xs[ys == top].mean())
    m[top - 9:top, xt] = 1                                   # vertical spike, 9 px
    left = xs.min()
    yl = int(ys[xs == left].mean())
    m[yl, left - 6:left] = 1                                 # horizontal spike
    if idx >= 1:
        for t in range(1, 8):                                # diagonal spike from the right end
            yy, xx = int(ys[xs == xs.max()].mean()) - t, xs.max() + t
            if 0 <= yy < H and 0 <= xx < W:
                m[yy, xx] = 1
    if idx == 2:
        m[40:43, 60:64] = 1                                  # a second, small component
        m[300, 20] = 1                                       # and an isolated pixel
    yy, xx = np.mgrid[:H, :W]
    depth = (0.45 + 1.5e-4 * (xx - cx) - 1.0e-4 * (yy - cy) + rng.normal(0, 1e-3, (H, W))).astype(np.float32)
    depth[m == 0] = 0.8
    return m, depth
