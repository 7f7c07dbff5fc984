"""Seeded synthetic frames for the grasp-selection path.

One definition shared by the oracle, the CUDA path, the CPU baseline and the
tests.  A frame is what the reference's two subscribers hand to
``LeafGraspNode.select_optimal_leaf`` (reference scripts/leaf_grasp_node_v3.py:185-205):

* ``labels``  int16 [H, W]  - 0 = background, leaf ids 1..N, one id per pixel
  (wire format msg/masks.msg: uint16[] imageData), later ids occlude earlier;
* ``depth``   float32 [H, W] - metres (msg/depth.msg: float32[] imageData);
* ``P``       float64 [3, 4] - CameraInfo.P; the path reads P[0,0], P[0,2],
  P[1,2] and P[0,3] (grasp_point_selector.py:145-150).

This is input generation, not product compute: it runs on the host with NumPy
and cv2.ellipse and is never timed.
"""
from __future__ import annotations

import dataclasses

import cv2
import numpy as np

# intrinsics of the reference's rig (scripts/leaf_grasp_node_2.py:24-27)
REF_F, REF_CX, REF_CY = 1750.68, 707.87, 494.07
REF_W, REF_H = 1440, 1080


@dataclasses.dataclass(frozen=True)
class FrameSpec:
    """Shape of one synthetic workload (BASELINE.json configs)."""

    height: int = REF_H
    width: int = REF_W
    n_leaves: int = 30
    # semi-axes in pixels; None = the SURVEY 8d figures scaled by width/1440
    a_range: tuple | None = None
    b_range: tuple | None = None
    margin: float | None = None

    @property
    def scale(self) -> float:
        return self.width / float(REF_W)

    def resolved(self):
        s = self.scale
        a = self.a_range or (90.0 * s, 180.0 * s)
        b = self.b_range or (50.0 * s, 100.0 * s)
        m = self.margin if self.margin is not None else 135.0 * s
        return a, b, m


CFG1 = FrameSpec(n_leaves=10)                                  # configs[0]
CFG2 = FrameSpec(n_leaves=30)                                  # configs[1] (the metric's workload)
CFG3 = FrameSpec(height=2160, width=3840, n_leaves=100)        # configs[2]
# small shapes for parity tests: leaves keep >= 10 000 px (leaf_scorer.py:80)
SMALL = FrameSpec(height=360, width=480, n_leaves=4, a_range=(80.0, 120.0),
                  b_range=(50.0, 70.0), margin=100.0)


def projection_matrix(spec: FrameSpec) -> np.ndarray:
    s = spec.scale
    f = REF_F * s
    P = np.zeros((3, 4), dtype=np.float64)
    P[0, 0] = f
    P[1, 1] = f
    P[0, 2] = REF_CX * s
    P[1, 2] = REF_CY * s
    P[2, 2] = 1.0
    P[0, 3] = -f * 0.1
    return P


def frame_seed(config_seed: int, frame_index: int) -> int:
    return int(config_seed) * 1_000_003 + int(frame_index)


def make_frame(spec: FrameSpec, config_seed: int, frame_index: int):
    """Return (labels int16 [H,W], depth float32 [H,W]) for one seeded frame."""
    H, W, N = spec.height, spec.width, spec.n_leaves
    (a_lo, a_hi), (b_lo, b_hi), margin = spec.resolved()
    rng = np.random.default_rng(frame_seed(config_seed, frame_index))
    labels = np.zeros((H, W), dtype=np.int16)
    depth = np.full((H, W), 0.8, dtype=np.float64)
    slope = 2e-4 / spec.scale
    stamp = np.zeros((H, W), dtype=np.uint8)
    for leaf in range(1, N + 1):
        cx = rng.uniform(margin, W - margin)
        cy = rng.uniform(margin, H - margin)
        a = rng.uniform(a_lo, a_hi)
        b = rng.uniform(b_lo, b_hi)
        ang = rng.uniform(0.0, 180.0)
        z0 = rng.uniform(0.3, 0.6)
        sx = rng.uniform(-slope, slope)
        sy = rng.uniform(-slope, slope)
        # paint into a window that certainly contains the ellipse (same pixels as a full-frame stamp)
        reach = int(round(a)) + 2
        x0, x1 = max(0, int(round(cx)) - reach), min(W, int(round(cx)) + reach + 1)
        y0, y1 = max(0, int(round(cy)) - reach), min(H, int(round(cy)) + reach + 1)
        stamp[y0:y1, x0:x1] = 0
        cv2.ellipse(stamp, (int(round(cx)), int(round(cy))), (int(round(a)), int(round(b))),
                    float(ang), 0.0, 360.0, 1, -1)
        sel = stamp[y0:y1, x0:x1].astype(bool)
        yy, xx = np.mgrid[y0:y1, x0:x1]
        labels[y0:y1, x0:x1][sel] = leaf
        plane = z0 + sx * (xx - cx) + sy * (yy - cy)
        depth[y0:y1, x0:x1][sel] = plane[sel]
    depth += rng.normal(0.0, 1e-3, size=(H, W))
    return labels, depth.astype(np.float32)


def make_batch(spec: FrameSpec, config_seed: int, first_index: int, count: int):
    """Stack ``count`` consecutive frames: labels [B,H,W] int16, depth [B,H,W] float32."""
    lab = np.empty((count, spec.height, spec.width), dtype=np.int16)
    dep = np.empty((count, spec.height, spec.width), dtype=np.float32)
    for i in range(count):
        lab[i], dep[i] = make_frame(spec, config_seed, first_index + i)
    return lab, dep


def seeded_state_dict(seed: int = 1234) -> dict:
    """Deterministic random-init GraspPointCNN weights for benchmarks and smoke runs (the reference ships no
    checkpoint): He-scaled conv / linear weights like model.py:89-100 and non-trivial BatchNorm statistics, keys as
    in SURVEY.md appendix A.9.  Draw for draw the tensors the test oracle builds for the same seed
    (tests/test_host_cpu.py checks the two stay equal), so benchmarked batches and golden vectors share one model."""
    import math

    import torch
    g = torch.Generator().manual_seed(seed)
    randn = lambda *shape: torch.randn(*shape, generator=g)
    sd = {}

    def batchnorm(prefix, n):
        sd[prefix + ".weight"] = 1.0 + 0.1 * randn(n)
        sd[prefix + ".bias"] = 0.1 * randn(n)
        sd[prefix + ".running_mean"] = 0.1 * randn(n)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(n, generator=g)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(100)

    cin = 9
    for blk, cout in enumerate((64, 128, 256)):
        for conv, ci in ((0, cin), (3, cout)):
            sd[f"encoder.{blk}.{conv}.weight"] = randn(cout, ci, 3, 3) * math.sqrt(2.0 / (cout * 9))
            sd[f"encoder.{blk}.{conv}.bias"] = randn(cout) * 0.05
            batchnorm(f"encoder.{blk}.{conv + 1}", cout)
        cin = cout
    sd["attention.0.weight"] = randn(1, 256, 1, 1) * math.sqrt(2.0) * 0.1
    sd["attention.0.bias"] = torch.zeros(1)
    widths = (256, 256, 128, 64, 1)
    for k, lin in enumerate((0, 4, 8, 12)):
        fan_in, fan_out = widths[k], widths[k + 1]
        sd[f"classifier.{lin}.weight"] = randn(fan_out, fan_in) * math.sqrt(2.0 / fan_in)
        sd[f"classifier.{lin}.bias"] = randn(fan_out) * 0.05
        if lin != 12:
            batchnorm(f"classifier.{lin + 1}", fan_out)
    return sd
