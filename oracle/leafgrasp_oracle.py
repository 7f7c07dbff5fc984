"""CPU oracle for the grasp-selection hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU with NumPy / OpenCV / SciPy / torch-CPU, the algorithm of the
reference's per-frame path so the CUDA implementation can be checked against it.  Nothing in the
product package imports it: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.

Reference files restated (paths relative to the reference checkout):
  scripts/utils/leaf_scorer.py:25-203,277-306          -> select_optimal_leaf
  scripts/utils/grasp_point_selector.py:256-288         -> score_maps / valid_regions
  scripts/utils/grasp_point_selector.py:502-701,718-752 -> the individual maps, leaf_orientation
  scripts/utils/grasp_point_selector.py:447-482         -> candidate_points
  scripts/utils/grasp_point_selector.py:59-143,392-445  -> extract_patch / patch_tensor / ml_score
  scripts/utils/grasp_point_selector.py:184-253         -> select_grasp_point
  scripts/utils/grasp_point_selector.py:152-180,754-826 -> grasp_point_3d / pre_grasp_point
  scripts/utils/image_processor.py:15-32,56-64          -> gaussian_kernel / smooth
  scripts/utils/ml_grasp_optimizer/model.py:5-128       -> cnn_forward

Parity pin: the reference ships no tests, fixtures or golden vectors for this path
(SURVEY.md section 4).  The pin is therefore the reference code itself, imported in the build
container by ``tests/golden/make_golden.py``; its outputs are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` holds this oracle to them.  One sub-result is PARITY UNPINNED: the
background pixel farthest from every leaf (leaf_scorer.py:69-71) comes from scikit-fmm 2022.3.26
(requirements.txt:5), a fast-marching solver that is not installed here and cannot be reproduced
bit for bit; the oracle (and the golden generator's shim) substitute the exact Euclidean distance
transform, whose arg-max the fast-marching field approximates.  paretoset 1.2.3 (requirements.txt:17)
is restated exactly (non-dominated rows, first of duplicates kept).

Two arithmetic modes:
  arith="reference"  the same NumPy / torch-CPU calls the reference makes (np.exp on float32,
                     cos(arctan2), F.conv2d) - used for the golden check and the CPU baseline;
  arith="strict"     every transcendental replaced by a correctly-rounded, platform-independent
                     definition (float32 exp := float32(exp(float64 x)); cos(arctan2(dy,dx)) := dx/r;
                     stencils accumulated tap by tap in float32 without fused multiply-add).  This is
                     the specification the CUDA path is held to bit-for-bit on integer results and
                     candidate indices.
OpenCV note: cv2's IPP float chamfer differs from its integer one; the oracle always runs with
``cv2.ipp.setUseIPP(False)``.
"""
from __future__ import annotations

import math

import cv2
import numpy as np
import scipy.ndimage as ndi
import torch
import torch.nn.functional as F

cv2.ipp.setUseIPP(False)

# ----------------------------------------------------------------------------------------------
# constants of the path (all from the reference, file:line beside each)
# ----------------------------------------------------------------------------------------------
MIN_LEAF_AREA = 10000          # leaf_scorer.py:80
DIST_SCALE_M = 0.3             # leaf_scorer.py:117
LEAF_WEIGHTS = (0.35, 0.35, 0.3)   # leaf_scorer.py:170
MIN_EDGE_DISTANCE = 20         # grasp_point_selector.py:25,285
OPTIMAL_EDGE_DISTANCE = 20     # grasp_point_selector.py:535
TOP_K = 20                     # grasp_point_selector.py:197
NMS_RADIUS = 10                # grasp_point_selector.py:198
PATCH = 32                     # grasp_point_selector.py:66
STEM_SE = 30                   # grasp_point_selector.py:696
PREGRASP_SE = 31               # grasp_point_selector.py:777-778
TRAD_WEIGHTS = (0.4, 0.3, 0.2, 0.1)           # approach, sdf_score, flatness, accessibility (grasp_point_selector.py:272-277)
CH_A, CH_B, CH_C = 65536, 91750, 143976     # OpenCV DIST_L2 5x5 weights 1, 1.4, 2.1969 in Q16
CH_DIST_MAX = 0xFFFFFFFF - CH_C             # OpenCV's saturation value (probed: all-ones image)
SCORE_CHANNELS = ("sdf_score", "approach_score", "flatness_map", "isolation_map",
                  "distance_map", "accessibility_map", "stem_penalty")   # grasp_point_selector.py:95-99


# ----------------------------------------------------------------------------------------------
# distance transforms
# ----------------------------------------------------------------------------------------------
def chamfer5_q16(mask: np.ndarray) -> np.ndarray:
    """Integer two-pass 5x5 chamfer transform, the arithmetic behind
    cv2.distanceTransform(mask, DIST_L2, 5) with IPP off (grasp_point_selector.py:266,529-530).

    Returns uint32 [H,W] in Q16.  Zero pixels of ``mask`` are the sources.  Written row by row:
    the 7 taps that reach into the two previous rows are a plain stencil, the in-row tap
    t[x] = min(t[x], t[x-1]+a) is a running minimum of (t[x] - a*x).
    """
    m = np.asarray(mask) != 0
    H, W = m.shape
    INF = np.int64(CH_DIST_MAX)
    a, b, c = CH_A, CH_B, CH_C
    t = np.full((H + 4, W + 4), INF, dtype=np.int64)
    xs = np.arange(W, dtype=np.int64) * a

    def sweep(rows, flip):
        for y in rows:
            r1 = t[y - 1] if not flip else t[y + 1]
            r2 = t[y - 2] if not flip else t[y + 2]
            cand = np.minimum.reduce([
                r2[1:W + 1] + c, r2[3:W + 3] + c,
                r1[0:W] + c, r1[1:W + 1] + b, r1[2:W + 2] + a, r1[3:W + 3] + b, r1[4:W + 4] + c,
            ])
            cur = t[y, 2:W + 2]
            u = np.minimum(cur, cand)
            if not flip:
                u = np.where(m[y - 2], u, 0)
                run = np.minimum.accumulate(np.minimum(u, INF) - xs) + xs
            else:
                ur = u[::-1]
                run = (np.minimum.accumulate(ur - xs) + xs)[::-1]
            t[y, 2:W + 2] = np.minimum(np.minimum(u, run), INF)

    sweep(range(2, H + 2), False)
    sweep(range(H + 1, 1, -1), True)
    return t[2:H + 2, 2:W + 2].astype(np.uint32)


def chamfer_norm_q16(dx, dy):
    """Closed form of the 5x5 chamfer norm: cost of the cheapest path made of a/b/c moves."""
    dx = np.abs(np.asarray(dx, dtype=np.int64))
    dy = np.abs(np.asarray(dy, dtype=np.int64))
    mx = np.maximum(dx, dy)
    mn = np.minimum(dx, dy)
    return np.where(2 * mn <= mx, CH_C * mn + CH_A * (mx - 2 * mn),
                    CH_C * (mx - mn) + CH_B * (2 * mn - mx))


def q16_to_float(q: np.ndarray) -> np.ndarray:
    """OpenCV's final conversion: float32(t) * float32(2^-16)."""
    return q.astype(np.float32) * np.float32(1.0 / 65536.0)


def chamfer5(mask_u8: np.ndarray, use_cv2: bool = True) -> np.ndarray:
    if use_cv2:
        return cv2.distanceTransform(np.ascontiguousarray(mask_u8, dtype=np.uint8), cv2.DIST_L2, 5)
    return q16_to_float(chamfer5_q16(mask_u8))


def edt_squared(nonzero_is_far: np.ndarray) -> np.ndarray:
    """Exact squared Euclidean distance (int64) from every non-zero pixel to the nearest zero pixel."""
    d = ndi.distance_transform_edt(np.asarray(nonzero_is_far) != 0)
    return np.rint(d * d).astype(np.int64)


def edt_squared_bruteforce(nonzero_is_far: np.ndarray) -> np.ndarray:
    """O(P*Z) restatement for tiny images, used to pin edt_squared / the CUDA EDT."""
    far = np.asarray(nonzero_is_far) != 0
    H, W = far.shape
    zy, zx = np.nonzero(~far)
    out = np.zeros((H, W), dtype=np.int64)
    if zy.size == 0:
        out[:] = np.iinfo(np.int32).max
        return out
    yy, xx = np.mgrid[0:H, 0:W]
    d2 = (yy[..., None] - zy) ** 2 + (xx[..., None] - zx) ** 2
    return d2.min(axis=-1).astype(np.int64)


# ----------------------------------------------------------------------------------------------
# morphology (OpenCV semantics restated; cv2 used as the fast path)
# ----------------------------------------------------------------------------------------------
def ellipse_se(n: int) -> np.ndarray:
    """cv2.getStructuringElement(MORPH_ELLIPSE, (n, n)) - taken from cv2 (OpenCV-version dependent,
    SURVEY.md hard part 5)."""
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (n, n))


def dilate_restated(mask_u8: np.ndarray, se: np.ndarray) -> np.ndarray:
    """cv2.dilate(mask, se) with the default anchor (n//2, n//2) and default border: a pixel is set
    iff any in-image pixel under the structuring element footprint is set."""
    m = np.asarray(mask_u8) != 0
    H, W = m.shape
    kh, kw = se.shape
    ay, ax = kh // 2, kw // 2
    out = np.zeros((H, W), dtype=bool)
    pad = np.zeros((H + kh, W + kw), dtype=bool)
    pad[ay:ay + H, ax:ax + W] = m
    for j in range(kh):
        for i in range(kw):
            if se[j, i]:
                out |= pad[j:j + H, i:i + W]
    return out.astype(np.uint8)


def dilate(mask_u8, se, use_cv2=True):
    if use_cv2:
        return cv2.dilate(np.ascontiguousarray(mask_u8, dtype=np.uint8), se)
    return dilate_restated(mask_u8, se)


# ----------------------------------------------------------------------------------------------
# orientation: largest outer contour -> minimum-area rectangle   (grasp_point_selector.py:718-752)
# ----------------------------------------------------------------------------------------------
def leaf_orientation(mask_u8: np.ndarray):
    """Returns (angle_rad or None, major, minor, center) exactly as the reference computes it."""
    contours, _ = cv2.findContours(np.ascontiguousarray(mask_u8, dtype=np.uint8),
                                   cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not contours:
        return None, None, None, None
    contour = max(contours, key=cv2.contourArea)
    (cx, cy), (w, h), ang = cv2.minAreaRect(contour)
    if w < h:
        ang = ang + 90
    return np.deg2rad(ang), max(w, h), min(w, h), (cx, cy)


def hull_from_points(pts: np.ndarray) -> np.ndarray:
    """Strictly convex hull of integer pixel coordinates, in the order cv2.convexHull(contour, clockwise=True)
    gives for a contour that starts at its raster-first pixel: start at the top-left pixel, down the left
    side, along the bottom, up the right side.  Built from the per-row extents with two monotone chains."""
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 2)
    ys = np.unique(pts[:, 1])
    lo = {int(y): int(pts[pts[:, 1] == y, 0].min()) for y in ys}
    hi = {int(y): int(pts[pts[:, 1] == y, 0].max()) for y in ys}

    def cross(a, b, c):
        return (b[0] - a[0]) * (c[1] - b[1]) - (b[1] - a[1]) * (c[0] - b[0])

    h = []
    for y in ys:
        p = (lo[int(y)], int(y))
        while len(h) >= 2 and cross(h[-2], h[-1], p) >= 0:
            h.pop()
        h.append(p)
    base = len(h)
    for y in ys[::-1]:
        p = (hi[int(y)], int(y))
        if h and h[-1] == p:
            continue
        while len(h) >= base + 1 and len(h) >= 2 and cross(h[-2], h[-1], p) >= 0:
            h.pop()
        h.append(p)
    if len(h) > 1 and h[-1] == h[0]:
        h.pop()
    while len(h) >= 3 and cross(h[-2], h[-1], h[0]) >= 0:
        h.pop()
    return np.array(h, dtype=np.int64)


def min_area_rect_restated(hull: np.ndarray):
    """OpenCV's rotatingCalipers(CALIPERS_MINAREARECT) and the tail of cv::minAreaRect restated in float32
    without fused multiply-add (strict-mode definition; cv2 4.13 agrees to the last bit on most inputs and to
    1-2 ulp on the rest).  Returns ((cx, cy), (w, h), angle_deg) as float32."""
    f32, f64 = np.float32, np.float64
    pts = np.asarray(hull, dtype=np.float32)
    n = len(pts)
    if n == 1:
        return (pts[0, 0], pts[0, 1]), (f32(0), f32(0)), f32(0)
    if n == 2:
        dx, dy = f64(pts[1, 0]) - f64(pts[0, 0]), f64(pts[1, 1]) - f64(pts[0, 1])
        return ((pts[0, 0] + pts[1, 0]) * f32(0.5), (pts[0, 1] + pts[1, 1]) * f32(0.5)), \
               (f32(np.sqrt(dx * dx + dy * dy)), f32(0)), f32(np.arctan2(dy, dx) * 180.0 / np.pi)
    vect = np.zeros((n, 2), f32)
    inv = np.zeros(n, f32)
    left = bottom = right = top = 0
    pt0 = pts[0].copy()
    left_x = right_x = pt0[0]
    top_y = bottom_y = pt0[1]
    for i in range(n):
        if pt0[0] < left_x:
            left_x, left = pt0[0], i
        if pt0[0] > right_x:
            right_x, right = pt0[0], i
        if pt0[1] > top_y:
            top_y, top = pt0[1], i
        if pt0[1] < bottom_y:
            bottom_y, bottom = pt0[1], i
        pt = pts[(i + 1) % n]
        dx, dy = f64(pt[0]) - f64(pt0[0]), f64(pt[1]) - f64(pt0[1])
        vect[i] = (f32(dx), f32(dy))
        inv[i] = f32(1.0 / np.sqrt(dx * dx + dy * dy))
        pt0 = pt.copy()
    orientation = f32(0)
    ax, ay = f64(vect[n - 1, 0]), f64(vect[n - 1, 1])
    for i in range(n):
        bx, by = f64(vect[i, 0]), f64(vect[i, 1])
        conv = ax * by - ay * bx
        if conv != 0:
            orientation = f32(1) if conv > 0 else f32(-1)
            break
        ax, ay = bx, by
    base_a, base_b = orientation, f32(0)
    seq = [bottom, right, top, left]
    minarea = f32(3.4028235e38)
    best = None
    for _ in range(n):
        dp = [base_a * vect[seq[0], 0] + base_b * vect[seq[0], 1],
              -base_b * vect[seq[1], 0] + base_a * vect[seq[1], 1],
              -base_a * vect[seq[2], 0] - base_b * vect[seq[2], 1],
              base_b * vect[seq[3], 0] - base_a * vect[seq[3], 1]]
        maxcos, main = dp[0] * inv[seq[0]], 0
        for i in range(1, 4):
            c = dp[i] * inv[seq[i]]
            if c > maxcos:
                main, maxcos = i, c
        p = seq[main]
        lx, ly = vect[p, 0] * inv[p], vect[p, 1] * inv[p]
        base_a, base_b = ((lx, ly), (ly, -lx), (-lx, -ly), (-ly, lx))[main]
        seq[main] = (seq[main] + 1) % n
        dx, dy = pts[seq[1], 0] - pts[seq[3], 0], pts[seq[1], 1] - pts[seq[3], 1]
        width = dx * base_a + dy * base_b
        dx, dy = pts[seq[2], 0] - pts[seq[0], 0], pts[seq[2], 1] - pts[seq[0], 1]
        height = -dx * base_b + dy * base_a
        area = width * height
        if area <= minarea:
            minarea = area
            best = (seq[3], base_a, width, base_b, height, seq[0])
    l, A1, w, B1, h, b = best
    A2, B2 = -B1, A1
    C1 = A1 * pts[l, 0] + pts[l, 1] * B1
    C2 = A2 * pts[b, 0] + pts[b, 1] * B2
    idet = f32(1) / (A1 * B2 - A2 * B1)
    px, py = (C1 * B2 - C2 * B1) * idet, (A1 * C2 - A2 * C1) * idet
    o2, o3, o4, o5 = A1 * w, B1 * w, A2 * h, B2 * h
    cx, cy = px + (o2 + o4) * f32(0.5), py + (o3 + o5) * f32(0.5)
    ww = f32(np.sqrt(f64(o2) * f64(o2) + f64(o3) * f64(o3)))
    hh = f32(np.sqrt(f64(o4) * f64(o4) + f64(o5) * f64(o5)))
    ang = f32(np.arctan2(f64(o3), f64(o2)) * 180.0 / np.pi)
    return (cx, cy), (ww, hh), ang


def leaf_orientation_restated(mask_u8: np.ndarray):
    """Strict-mode orientation: cv2 picks the outer contour of largest polygon area; its hull and the
    rectangle come from the restatements above."""
    contours, _ = cv2.findContours(np.ascontiguousarray(mask_u8, dtype=np.uint8),
                                   cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not contours:
        return None, None, None, None
    contour = max(contours, key=cv2.contourArea)
    (cx, cy), (w, h), ang = min_area_rect_restated(hull_from_points(contour.reshape(-1, 2)))
    ang = float(ang)
    if w < h:
        ang = ang + 90
    return np.deg2rad(ang), float(max(w, h)), float(min(w, h)), (float(cx), float(cy))


def moore_contour_area(mask_u8: np.ndarray, sx: int, sy: int) -> float:
    """cv2.contourArea of the outer border of the 8-connected component whose raster-first pixel is
    (sx, sy), by Moore-neighbour tracing and the shoelace formula (the rule csrc/lg_orient.cu uses)."""
    DX = (1, 1, 0, -1, -1, -1, 0, 1)
    DY = (0, 1, 1, 1, 0, -1, -1, -1)
    H, W = mask_u8.shape
    bit = lambda x, y: 0 <= x < W and 0 <= y < H and mask_u8[y, x] != 0
    cx, cy, db, first, a00 = sx, sy, 4, -1, 0
    while True:
        d = -1
        for k in range(1, 9):
            dd = (db + k) % 8
            if bit(cx + DX[dd], cy + DY[dd]):
                d = dd
                break
        if d < 0 or (cx == sx and cy == sy and first >= 0 and d == first):
            break
        if first < 0:
            first = d
        nx, ny = cx + DX[d], cy + DY[d]
        a00 += cx * ny - nx * cy
        cx, cy = nx, ny
        db = (d + (5 if d % 2 else 6)) % 8
    return abs(a00) * 0.5


# ----------------------------------------------------------------------------------------------
# stage 1: optimal leaf            (leaf_scorer.py:25-203, 277-306)
# ----------------------------------------------------------------------------------------------
def pareto_front_max(scores: np.ndarray) -> np.ndarray:
    """paretoset(scores, sense=['max']*k) restated: keep rows no other row dominates; of exact
    duplicates keep the first (paretoset's distinct=True default)."""
    n = scores.shape[0]
    keep = np.ones(n, dtype=bool)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            ge = np.all(scores[j] >= scores[i])
            gt = np.any(scores[j] > scores[i])
            if ge and gt:
                keep[i] = False
                break
            if ge and not gt and j < i:      # exact duplicate seen earlier
                keep[i] = False
                break
    return keep


def visibility_score(leaf_mask: np.ndarray) -> float:
    h, w = leaf_mask.shape
    ys, xs = np.where(leaf_mask)
    if len(ys) == 0:
        return 0.0
    touching = (np.sum(leaf_mask[0, :]) + np.sum(leaf_mask[-1, :]) +
                np.sum(leaf_mask[:, 0]) + np.sum(leaf_mask[:, -1]))
    if touching > 0:
        return 0.0
    mx, my = np.mean(xs), np.mean(ys)
    half_w, half_h = w / 2, h / 2
    return 1.0 - (np.sqrt((mx - half_w) ** 2 + (my - half_h) ** 2) / np.sqrt(half_w ** 2 + half_h ** 2))


def clutter_extrema(labels: np.ndarray):
    """(row, col) of the field minimum and maximum used by the clutter score (leaf_scorer.py:67-71).
    The field is 0 on every leaf pixel and the distance to the nearest leaf pixel elsewhere, so the
    arg-min is the first leaf pixel in raster order; the arg-max is taken on the exact EDT
    (scikit-fmm substitute - parity unpinned, see the module header)."""
    leafy = labels >= 1
    field = ndi.distance_transform_edt(~leafy)
    pmin = np.unravel_index(field.argmin(), field.shape)
    pmax = np.unravel_index(field.argmax(), field.shape)
    return (int(pmin[0]), int(pmin[1])), (int(pmax[0]), int(pmax[1]))


def select_optimal_leaf(labels: np.ndarray, depth: np.ndarray, f: float, cx: float, cy: float):
    """Returns dict(leaf_id=int|None, tall=[ids], candidates=[...], pmin, pmax, medians={id: f32})."""
    labels = np.asarray(labels)
    depth = np.asarray(depth, dtype=np.float32)
    ids = np.unique(labels)[1:]            # torch.unique(...)[1:] - drops the smallest id present
    out = dict(leaf_id=None, tall=[], candidates=[], pmin=None, pmax=None, medians={})
    masks, medians = [], []
    for i in ids:
        m = labels == i
        masks.append(m)
        vals = depth[m]
        if len(vals) > 0:
            medians.append(np.median(vals))
    if not medians:
        return out
    med = np.array(medians)
    mean_of_medians = np.mean(med)
    tall = [int(ids[k]) for k, d in enumerate(medians) if d < mean_of_medians]
    out["tall"] = tall
    out["medians"] = {int(ids[k]): medians[k] for k in range(len(medians))}
    pmin, pmax = clutter_extrema(labels)
    out["pmin"], out["pmax"] = pmin, pmax
    cands = []
    for k, i in enumerate(ids):
        m = masks[k]
        area = np.sum(m)
        if area < MIN_LEAF_AREA:
            continue
        ys, xs = np.where(m)
        c = (np.mean(xs), np.mean(ys))
        dmin = np.sqrt((c[0] - pmin[1]) ** 2 + (c[1] - pmin[0]) ** 2)
        dmax = np.sqrt((c[0] - pmax[1]) ** 2 + (c[1] - pmax[0]) ** 2)
        tot = dmin + dmax
        clutter = dmin / tot if tot > 0 else 0
        md = np.mean(depth[m])                     # float32 scalar
        X = (md * (xs - cx)) / f
        Y = (md * (ys - cy)) / f
        Z = np.full_like(X, md)
        mean_dist = np.mean(np.sqrt(X ** 2 + Y ** 2 + Z ** 2))
        dist_score = np.exp(-mean_dist / DIST_SCALE_M)
        vis = visibility_score(m)
        cands.append(dict(leaf_id=int(i), scores=np.array([clutter, dist_score, vis], dtype=np.float64),
                          is_tall=int(i) in tall, area=int(area), centroid=c, mean_depth=md,
                          mean_distance=mean_dist))
    out["candidates"] = cands
    if not cands:
        return out
    tall_c = [c for c in cands if c["is_tall"]]
    group = tall_c if tall_c else cands
    sc = np.stack([c["scores"] for c in group])
    if tall_c:
        sc = sc * 1.1
    front = pareto_front_max(sc)
    pool = [c for k, c in enumerate(group) if front[k]] or group
    w = np.array(LEAF_WEIGHTS)
    best, best_score = None, float("-inf")
    for c in pool:
        s = np.sum(w * c["scores"])
        if s > best_score:
            best_score, best = s, c["leaf_id"]
    out["leaf_id"] = best
    return out


# ----------------------------------------------------------------------------------------------
# stage 2: per-pixel score maps        (grasp_point_selector.py:256-288, 502-701)
# ----------------------------------------------------------------------------------------------
def gaussian_kernel(size: int = 5) -> np.ndarray:
    """image_processor.py:25-32 - sigma = size/6, normalised in float64, stored float32."""
    sigma = size / 6.0
    c = size // 2
    x, y = np.meshgrid(np.arange(size), np.arange(size))
    k = np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2))
    return (k / k.sum()).astype(np.float32)


SOBEL_X = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=np.float32)   # image_processor.py:19
SOBEL_Y = SOBEL_X.T.copy()


def _exp32(x32: np.ndarray, arith: str) -> np.ndarray:
    if arith == "reference":
        return np.exp(x32)
    return np.exp(x32.astype(np.float64)).astype(np.float32)


def _corr_taps(padded: np.ndarray, k: np.ndarray, H: int, W: int) -> np.ndarray:
    """Cross-correlation accumulated tap by tap (row-major over the kernel) in float32 with separate
    rounding of each product and each sum; zero-weight taps are skipped.  Strict-mode stencil."""
    acc = None
    kh, kw = k.shape
    for j in range(kh):
        for i in range(kw):
            if k[j, i] == 0:
                continue
            term = (k[j, i] * padded[j:j + H, i:i + W]).astype(np.float32)
            acc = term if acc is None else (acc + term).astype(np.float32)
    return acc


def smooth_depth(depth_patch: np.ndarray, arith: str = "reference") -> np.ndarray:
    """ImageProcessor.smooth_depth (image_processor.py:56-64): reflect padding by 2, then the 5x5 Gaussian; float32 [h,w].
    "reference" makes the reference's torch calls, "strict" accumulates the 25 taps one by one (what the CUDA kernel does)."""
    z = np.ascontiguousarray(depth_patch, dtype=np.float32)
    g = gaussian_kernel(5)
    if arith == "reference":
        pz = F.pad(torch.from_numpy(z)[None, None], (2, 2, 2, 2), mode="reflect")
        return F.conv2d(pz, torch.from_numpy(g)[None, None]).squeeze().numpy()
    return _corr_taps(np.pad(z, 2, mode="reflect"), g, z.shape[0], z.shape[1])


def flatness_map(depth: np.ndarray, mask_u8: np.ndarray, arith: str = "reference") -> np.ndarray:
    """grasp_point_selector.py:262,635-657 + image_processor.py:56-64.  float32 [H,W], NOT masked."""
    H, W = mask_u8.shape
    z = (np.asarray(depth, dtype=np.float32) * mask_u8.astype(np.float32)).astype(np.float32)
    g = gaussian_kernel(5)
    if arith == "reference":
        zt = torch.from_numpy(z)
        pz = F.pad(zt[None, None], (2, 2, 2, 2), mode="reflect")
        s = F.conv2d(pz, torch.from_numpy(g)[None, None]).squeeze()
        ps = F.pad(s[None, None], (1, 1, 1, 1), mode="reflect")
        dx = F.conv2d(ps, torch.from_numpy(SOBEL_X)[None, None]).squeeze()
        dy = F.conv2d(ps, torch.from_numpy(SOBEL_Y)[None, None]).squeeze()
        mag = torch.sqrt(dx ** 2 + dy ** 2)
        return torch.exp(-mag * 5).numpy()
    pz = np.pad(z, 2, mode="reflect")
    s = _corr_taps(pz, g, H, W)
    ps = np.pad(s, 1, mode="reflect")
    dx = _corr_taps(ps, SOBEL_X, H, W)
    dy = _corr_taps(ps, SOBEL_Y, H, W)
    mag = np.sqrt((dx * dx).astype(np.float32) + (dy * dy).astype(np.float32)).astype(np.float32)
    return _exp32((-mag * np.float32(5)).astype(np.float32), arith)


def sdf_score_map(mask_u8, cx, cy, arith="reference", use_cv2=True, want_parts=False):
    """grasp_point_selector.py:526-567."""
    di = chamfer5(mask_u8, use_cv2)
    do = chamfer5(1 - mask_u8, use_cv2)
    sdf = di - do
    interior = _exp32(-((di - OPTIMAL_EDGE_DISTANCE) ** 2) / (2 * OPTIMAL_EDGE_DISTANCE ** 2), arith)
    sdf_max = np.max(np.abs(sdf))
    sdf = sdf / sdf_max
    H, W = mask_u8.shape
    ys, xs = np.indices((H, W))
    vx = xs - cx
    vy = ys - cy
    nrm = np.sqrt(vx * vx + vy * vy)
    nrm[nrm == 0] = 1
    vx = vx / nrm
    vy = vy / nrm
    angle, _, _, _ = leaf_orientation(mask_u8) if arith == "reference" else leaf_orientation_restated(mask_u8)
    if angle is not None:
        ca, sa = np.cos(angle), np.sin(angle)
        align = np.abs(vx * sa - vy * ca)
    else:
        align = np.ones_like(sdf)
    score = (0.4 * interior + 0.4 * align + 0.2 * sdf) * mask_u8
    if want_parts:
        return score, dict(di=di, do=do, sdf_max=sdf_max, angle=angle)
    return score


def approach_score_map(mask_u8, f, cx, cy):
    """grasp_point_selector.py:569-593 - cosine between the pixel ray and the optical axis."""
    H, W = mask_u8.shape
    ys, xs = np.indices((H, W))
    vx = xs - cx
    vy = ys - cy
    nrm = np.sqrt(vx * vx + vy * vy + np.full((H, W), f) * np.full((H, W), f))
    nrm[nrm == 0] = 1
    return np.abs(np.full((H, W), f) / nrm) * mask_u8


def accessibility_map(mask_u8, cx, cy, arith="reference"):
    """grasp_point_selector.py:502-524."""
    H, W = mask_u8.shape
    yg, xg = np.ogrid[:H, :W]
    r = np.sqrt((xg - cx) ** 2 + (yg - cy) ** 2)
    near = 1 - (r / np.sqrt(W ** 2 + H ** 2))
    if arith == "reference":
        fwd = np.cos(np.arctan2(yg - cy, xg - cx))
    else:
        rr = np.where(r == 0, 1.0, r)
        fwd = np.where(r == 0, 1.0, (xg - cx) / rr)
    return (0.7 * near + 0.3 * fwd) * mask_u8


def isolation_map(mask_u8):
    """Effective result of grasp_point_selector.py:595-633.  The routine is handed the single-leaf
    binary mask, so its "other leaves" image is identically zero, both chamfer transforms act on an
    all-ones image, both normalised fields are 1 up to the 1e-6 guard, and what is left is the
    row ramp linspace(1.0, 0.2, H) on the leaf (pinned in the golden check)."""
    H, W = mask_u8.shape
    unit = np.float32(0.7) * np.float32(1.0) + np.float32(0.3) * np.float32(1.0)   # == 1.0f
    ramp = np.linspace(1.0, 0.2, H)[:, np.newaxis]
    return (np.float64(unit) * np.tile(ramp, (1, W))) * mask_u8.astype(np.uint8)


def stem_penalty_map(mask_u8, use_cv2=True):
    """grasp_point_selector.py:688-701."""
    H, W = mask_u8.shape
    bottom = np.zeros_like(mask_u8)
    third = H // 3
    bottom[-third:, :] = 1
    se = ellipse_se(STEM_SE)
    return (dilate(mask_u8 & bottom, se, use_cv2) & mask_u8).astype(np.float32)


def score_maps(mask_u8, depth, f, cx, cy, arith="reference", use_cv2=True):
    """grasp_point_selector.py:256-280.  Returns the reference's dict of 8 maps."""
    mask_u8 = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    sdf_score, parts = sdf_score_map(mask_u8, cx, cy, arith, use_cv2, want_parts=True)
    s = {
        "sdf_score": sdf_score,
        "approach_score": approach_score_map(mask_u8, f, cx, cy),
        "flatness_map": flatness_map(depth, mask_u8, arith),
        "isolation_map": isolation_map(mask_u8),
        "distance_map": parts["di"],
        "accessibility_map": accessibility_map(mask_u8, cx, cy, arith),
        "stem_penalty": stem_penalty_map(mask_u8, use_cv2),
    }
    wa, ws, wf, wc = TRAD_WEIGHTS      # module-level so that a test can restate the README's set (SURVEY.md 8a')
    s["traditional_score"] = (wa * s["approach_score"] + ws * s["sdf_score"] +
                              wf * s["flatness_map"] + wc * s["accessibility_map"]) * (1 - s["stem_penalty"])
    s["_parts"] = parts
    return s


def valid_regions(mask_u8, scores):
    """grasp_point_selector.py:282-288."""
    return (scores["distance_map"] > MIN_EDGE_DISTANCE) & (mask_u8 > 0) & (scores["stem_penalty"] < 0.8)


# ----------------------------------------------------------------------------------------------
# candidates            (grasp_point_selector.py:447-482)
# ----------------------------------------------------------------------------------------------
def candidate_points(score_map, valid, top_k=TOP_K, min_distance=NMS_RADIUS):
    """Greedy pick in descending key order with a +-min_distance mark around every pick.
    Ties are defined as flat index descending (a stable ascending sort read backwards); the
    reference's np.argsort is unstable, so only inputs with distinct keys pin the order."""
    key = score_map * valid
    H, W = key.shape
    order = np.argsort(key.ravel(), kind="stable")[::-1]
    used = np.zeros((H, W), dtype=bool)
    picks = []
    for idx in order:
        if len(picks) >= top_k:
            break
        y, x = divmod(int(idx), W)
        y0, y1 = max(0, y - min_distance), min(y + min_distance + 1, H)
        x0, x1 = max(0, x - min_distance), min(x + min_distance + 1, W)
        if not used[y0:y1, x0:x1].any():
            picks.append((x, y))
            used[y0:y1, x0:x1] = True
    return picks


# ----------------------------------------------------------------------------------------------
# patches + CNN         (grasp_point_selector.py:59-143, 392-445; model.py:5-128)
# ----------------------------------------------------------------------------------------------
def extract_patch(arr, x, y, size=PATCH):
    """Window rows [y-16, y+16), cols [x-16, x+16) with edge replication."""
    h = size // 2
    H, W = arr.shape
    ys = np.clip(np.arange(y - h, y + h), 0, H - 1)
    xs = np.clip(np.arange(x - h, x + h), 0, W - 1)
    return arr[np.ix_(ys, xs)]


def patch_needs_padding(x, y, H, W, size=PATCH):
    h = size // 2
    return x - h < 0 or y - h < 0 or x + h > W or y + h > H


def _minmax32(p):
    p = p.astype(np.float32)
    lo, hi = p.min(), p.max()
    if hi > lo:
        p = ((p - lo) / (hi - lo)).astype(np.float32)
    return p


def patch_tensor(mask_u8, depth, scores, x, y):
    """float32 [9,32,32]: depth, mask, then the 7 score maps in SCORE_CHANNELS order; every channel
    except the mask is min-max normalised over the patch when max > min."""
    ch = [_minmax32(extract_patch(np.asarray(depth, dtype=np.float32), x, y)),
          extract_patch(mask_u8, x, y).astype(np.float32)]
    for name in SCORE_CHANNELS:
        ch.append(_minmax32(extract_patch(np.asarray(scores[name]), x, y)))
    return np.stack(ch).astype(np.float32)


def cnn_forward(sd: dict, x: torch.Tensor, n_blocks: int = 3) -> torch.Tensor:
    """GraspPointCNN.forward in eval mode (model.py:102-128), spatial attention, from a state_dict."""
    eps = 1e-5
    for b in range(n_blocks):
        for conv, bn in ((0, 1), (3, 4)):
            x = F.conv2d(x, sd[f"encoder.{b}.{conv}.weight"], sd[f"encoder.{b}.{conv}.bias"], padding=1)
            x = F.batch_norm(x, sd[f"encoder.{b}.{bn}.running_mean"], sd[f"encoder.{b}.{bn}.running_var"],
                             sd[f"encoder.{b}.{bn}.weight"], sd[f"encoder.{b}.{bn}.bias"], False, 0.0, eps)
            x = F.relu(x)
        x = F.max_pool2d(x, 2)
    att = torch.sigmoid(F.conv2d(x, sd["attention.0.weight"], sd["attention.0.bias"]))
    x = (x * att).mean(dim=(2, 3))
    for lin, bn in ((0, 1), (4, 5), (8, 9)):
        x = F.linear(x, sd[f"classifier.{lin}.weight"], sd[f"classifier.{lin}.bias"])
        x = F.batch_norm(x, sd[f"classifier.{bn}.running_mean"], sd[f"classifier.{bn}.running_var"],
                         sd[f"classifier.{bn}.weight"], sd[f"classifier.{bn}.bias"], False, 0.0, eps)
        x = F.relu(x)
    return F.linear(x, sd["classifier.12.weight"], sd["classifier.12.bias"])


def ml_rescale(logit: float) -> float:
    """sigmoid then tanh(3 s)/2 + 1/2     (grasp_point_selector.py:133-136)."""
    s = torch.sigmoid(torch.tensor(logit, dtype=torch.float32)).item()
    return float(np.tanh(s * 3.0) * 0.5 + 0.5)


def fuse(picks, trad_at, ml_scores):
    """grasp_point_selector.py:205-237.  ml_scores[i] is None where the reference gets no ML score."""
    best, best_score, ml_used = picks[0], trad_at[0], False
    if len(picks) > 1:
        for p, t, ml in zip(picks, trad_at, ml_scores):
            if ml is None:
                continue
            conf = 1.0 - abs(ml - 0.5) * 2
            w = min(0.3, conf * 0.6)
            comb = (1.0 - w) * t + w * ml
            if comb > best_score:
                best_score, best, ml_used = comb, p, True
    return best, best_score, ml_used


# ----------------------------------------------------------------------------------------------
# 3-D points            (grasp_point_selector.py:152-180, 754-826)
# ----------------------------------------------------------------------------------------------
def grasp_point_3d(pt, depth, f, cx, cy):
    u, v = pt
    z = float(depth[v, u])
    return ((z * (u - cx)) / f, (z * (v - cy)) / f, z)


def pre_grasp_point(p3, mask_u8, f, cx, cy, use_cv2=True):
    g = np.array(p3)
    d = g / np.linalg.norm(g)
    blocked = dilate(mask_u8, ellipse_se(PREGRASP_SE), use_cv2)
    H, W = mask_u8.shape
    for dist in np.arange(0.05, 0.10, 0.01):
        t = (p3[0] - d[0] * dist, p3[1] - d[1] * dist, p3[2])
        u = int((t[0] * f / t[2]) + cx)
        v = int((t[1] * f / t[2]) + cy)
        if not (0 <= u < W and 0 <= v < H):
            continue
        if blocked[v, u] == 0:
            if np.linalg.norm(np.array(t) - np.array(p3)) >= 0.05:
                return t
    return (p3[0] - d[0] * 0.10, p3[1] - d[1] * 0.10, p3[2])


# ----------------------------------------------------------------------------------------------
# whole stages
# ----------------------------------------------------------------------------------------------
def select_grasp_point(mask_u8, depth, f, cx, cy, state_dict=None, arith="reference", use_cv2=True,
                       want_debug=False):
    """grasp_point_selector.py:184-253.  Returns ((x,y),(X,Y,Z),(X,Y,Z)) or (None,None,None)."""
    mask_u8 = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    H, W = mask_u8.shape
    s = score_maps(mask_u8, depth, f, cx, cy, arith, use_cv2)
    valid = valid_regions(mask_u8, s)
    picks = candidate_points(s["traditional_score"], valid)
    if not picks:
        return (None, None, None) if not want_debug else ((None, None, None), {})
    trad_at = [s["traditional_score"][y, x] for (x, y) in picks]
    ml, logits = [None] * len(picks), [None] * len(picks)
    if state_dict is not None and len(picks) > 1:
        for i, (x, y) in enumerate(picks):
            # the reference hands a bool tensor to replicate padding, which torch rejects, so a
            # candidate whose window leaves the image gets no ML score
            if patch_needs_padding(x, y, H, W):
                continue
            pt = torch.from_numpy(patch_tensor(mask_u8, depth, s, x, y))[None]
            with torch.no_grad():
                lg = cnn_forward(state_dict, pt).item()
            logits[i] = lg
            ml[i] = ml_rescale(lg)
    best, best_score, ml_used = fuse(picks, trad_at, ml)
    p3 = grasp_point_3d(best, depth, f, cx, cy)
    pre = pre_grasp_point(p3, mask_u8, f, cx, cy, use_cv2)
    res = (best, p3, pre)
    if want_debug:
        return res, dict(scores=s, valid=valid, picks=picks, trad_at=trad_at, ml=ml, logits=logits,
                         best_score=best_score, ml_used=ml_used)
    return res


def process_frame(labels, depth, P, state_dict=None, arith="reference", use_cv2=True):
    """leaf_grasp_node_v3.py:102-133 without ROS: leaf selection, then grasp selection on that leaf."""
    f, cx, cy = P[0, 0], P[0, 2], P[1, 2]
    sel = select_optimal_leaf(labels, depth, f, cx, cy)
    if sel["leaf_id"] is None:
        return dict(leaf_id=None, grasp=(None, None, None), debug=None, leaf=sel)
    mask = (np.asarray(labels) == sel["leaf_id"]).astype(np.uint8)
    res, dbg = select_grasp_point(mask, depth, f, cx, cy, state_dict, arith, use_cv2, want_debug=True)
    return dict(leaf_id=sel["leaf_id"], grasp=res, debug=dbg, leaf=sel)


def seeded_state_dict(seed: int = 1234) -> dict:
    """A deterministic GraspPointCNN state_dict (the reference ships no checkpoint, .gitignore:16):
    Kaiming-style weights like model.py:89-100, plus non-trivial BatchNorm statistics so that the
    folded path is exercised.  Keys follow SURVEY.md appendix A.9."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 9
    for b, cout in enumerate((64, 128, 256)):
        for conv, bn, ci in ((0, 1, cin), (3, 4, cout)):
            sd[f"encoder.{b}.{conv}.weight"] = torch.randn(cout, ci, 3, 3, generator=g) * math.sqrt(2.0 / (cout * 9))
            sd[f"encoder.{b}.{conv}.bias"] = torch.randn(cout, generator=g) * 0.05
            sd[f"encoder.{b}.{bn}.weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sd[f"encoder.{b}.{bn}.bias"] = 0.1 * torch.randn(cout, generator=g)
            sd[f"encoder.{b}.{bn}.running_mean"] = 0.1 * torch.randn(cout, generator=g)
            sd[f"encoder.{b}.{bn}.running_var"] = 0.5 + torch.rand(cout, generator=g)
            sd[f"encoder.{b}.{bn}.num_batches_tracked"] = torch.tensor(100)
        cin = cout
    sd["attention.0.weight"] = torch.randn(1, 256, 1, 1, generator=g) * math.sqrt(2.0 / 1)  * 0.1
    sd["attention.0.bias"] = torch.zeros(1)
    dims = (256, 256, 128, 64, 1)
    for k, lin in enumerate((0, 4, 8, 12)):
        i, o = dims[k], dims[k + 1]
        sd[f"classifier.{lin}.weight"] = torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)
        sd[f"classifier.{lin}.bias"] = torch.randn(o, generator=g) * 0.05
        if lin != 12:
            bn = lin + 1
            sd[f"classifier.{bn}.weight"] = 1.0 + 0.1 * torch.randn(o, generator=g)
            sd[f"classifier.{bn}.bias"] = 0.1 * torch.randn(o, generator=g)
            sd[f"classifier.{bn}.running_mean"] = 0.1 * torch.randn(o, generator=g)
            sd[f"classifier.{bn}.running_var"] = 0.5 + torch.rand(o, generator=g)
            sd[f"classifier.{bn}.num_batches_tracked"] = torch.tensor(100)
    return sd


def seeded_state_dict_from_shapes(shapes: dict, seed: int) -> dict:
    """Deterministic weights for ANY GraspPointCNN architecture, from its state_dict key -> shape table (keys are
    visited in sorted order).  Used for the architecture-variant golden vectors (tests/golden/make_cnn_variants.py):
    the reference ships no checkpoints, so generator and tests rebuild the same tensors from the same table."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        elif len(shp) == 1:                      # BatchNorm weight
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:                                    # conv / linear weight
            fan_in = int(np.prod(shp[1:]))
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in)
    return sd


# ----------------------------------------------------------------------------------------------
# training-sample collector      (ml_grasp_optimizer/data_collector.py:83-348, 420-487)
# ----------------------------------------------------------------------------------------------
# The reference draws its negatives, its depth noise and its score jitter from the unseeded global
# generators of `random` and `torch`.  Those streams cannot be pinned, so the restatement takes the
# generator as an argument: `CollectorRng` below is the counter-based one the CUDA path uses, and
# tests/golden/make_collector.py injects the very same object into the unmodified reference module
# (as its `random` / `torch.randn_like`), so every other step of the reference is held bit for bit.
COLLECTOR_NEG_MAX = 3          # data_collector.py:306
COLLECTOR_ATTEMPTS = 10        # data_collector.py:303
KIND_POSITIVE, KIND_ROT90, KIND_ROT180, KIND_ROT270, KIND_TIP, KIND_STEM, KIND_EDGE = range(7)
_M64 = (1 << 64) - 1
# streams of the counter-based generator
RNG_NOISE_FACTOR, RNG_SCORE_JITTER, RNG_PICK, RNG_NORMAL = 1, 2, 3, 4


def mix64(z: int) -> int:
    """splitmix64 finaliser."""
    z &= _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _mix64_np(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


class CollectorRng:
    """draw(stream, a, b) = mix64(mix64(seed + 0x9E3779B97F4A7C15 * (frame + 1)) ^ (stream << 56 | a << 32 | b))."""

    def __init__(self, seed: int, frame: int):
        self.base = mix64(seed + 0x9E3779B97F4A7C15 * (frame + 1))

    def draw(self, stream: int, a: int, b: int) -> int:
        return mix64(self.base ^ ((stream << 56) | (a << 32) | b))

    def uniform01(self, stream: int, a: int, b: int = 0) -> float:
        return (self.draw(stream, a, b) >> 11) * 2.0 ** -53

    def noise_factor(self, k: int) -> float:          # random.uniform(0.01, 0.02), data_collector.py:271
        return 0.01 + (0.02 - 0.01) * self.uniform01(RNG_NOISE_FACTOR, k)

    def score_jitter(self, k: int) -> float:          # random.uniform(0.95, 1.0), data_collector.py:281
        return 0.95 + (1.0 - 0.95) * self.uniform01(RNG_SCORE_JITTER, k)

    def pick(self, attempt: int, kind: int, n: int) -> int:      # random.sample(points, 1), :319-327
        return self.draw(RNG_PICK, attempt, kind) % n

    def normal_patch(self, k: int) -> np.ndarray:
        """float32 [32,32] standard normals for rotation k (Box-Muller in float64, one pair per pixel)."""
        i = np.arange(PATCH * PATCH, dtype=np.uint64)
        with np.errstate(over="ignore"):
            key = np.uint64(self.base) ^ ((np.uint64(RNG_NORMAL) << np.uint64(56)) | (np.uint64(k) << np.uint64(32)))
            a = _mix64_np(key ^ (i * np.uint64(2)))
            b = _mix64_np(key ^ (i * np.uint64(2) + np.uint64(1)))
        u1 = ((a >> np.uint64(11)).astype(np.float64) + 1.0) * 2.0 ** -53
        u2 = (b >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        return z.astype(np.float32).reshape(PATCH, PATCH)


def collector_tip_points(mask_u8: np.ndarray, use_cv2: bool = True):
    """data_collector.py:420-440: local maxima (5x5) of the chamfer field on the leaf, largest quarter.
    list.sort is stable, so ties keep np.where's raster order."""
    mask_u8 = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    dist = chamfer5(mask_u8, use_cv2)
    if use_cv2:
        dil = cv2.dilate(dist, np.ones((5, 5), np.uint8))
    else:
        dil = ndi.maximum_filter(dist, size=5, mode="constant", cval=-np.inf)
    ys, xs = np.where((dil == dist) & (mask_u8 > 0))
    pts = list(zip(xs.tolist(), ys.tolist()))
    pts.sort(key=lambda p: dist[p[1], p[0]], reverse=True)
    return pts[:max(1, len(pts) // 4)]


def erode5_twice(mask_u8: np.ndarray, use_cv2: bool = True) -> np.ndarray:
    """cv2.erode(m, ellipse 5x5, iterations=2): two literal passes, out-of-image pixels never constrain."""
    se = ellipse_se(5)
    if use_cv2:
        return cv2.erode(np.ascontiguousarray(mask_u8, dtype=np.uint8), se, iterations=2)
    m = np.asarray(mask_u8) != 0
    H, W = m.shape
    for _ in range(2):
        pad = np.ones((H + 4, W + 4), dtype=bool)
        pad[2:2 + H, 2:2 + W] = m
        out = np.ones((H, W), dtype=bool)
        for j in range(5):
            for i in range(5):
                if se[j, i]:
                    out &= pad[j:j + H, i:i + W]
        m = out
    return m.astype(np.uint8)


def collector_stem_points(mask_u8: np.ndarray, use_cv2: bool = True):
    """data_collector.py:442-459: leaf pixels of the bottom quarter of the image that survive two erosions."""
    H = mask_u8.shape[0]
    stem = np.array(mask_u8, dtype=np.uint8, copy=True)
    stem[:int(0.75 * H)] = 0
    ys, xs = np.where(erode5_twice(stem, use_cv2) > 0)
    return list(zip(xs.tolist(), ys.tolist()))


def outer_border(mask_u8: np.ndarray, sx: int, sy: int):
    """Pixels of the outer border of the component whose raster-first pixel is (sx, sy), in the order
    cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_NONE) lists them: the Moore trace of
    moore_contour_area runs the other way round, so it is the start followed by that trace reversed."""
    DX = (1, 1, 0, -1, -1, -1, 0, 1)
    DY = (0, 1, 1, 1, 0, -1, -1, -1)
    H, W = mask_u8.shape
    bit = lambda x, y: 0 <= x < W and 0 <= y < H and mask_u8[y, x] != 0
    cx, cy, db, first, pts = sx, sy, 4, -1, []
    while True:
        d = -1
        for k in range(1, 9):
            dd = (db + k) % 8
            if bit(cx + DX[dd], cy + DY[dd]):
                d = dd
                break
        if d < 0:
            pts.append((cx, cy))
            break
        if cx == sx and cy == sy and first >= 0 and d == first:
            break
        if first < 0:
            first = d
        pts.append((cx, cy))
        cx, cy = cx + DX[d], cy + DY[d]
        db = (d + (5 if d % 2 else 6)) % 8
    return [pts[0]] + pts[:0:-1]


def collector_edge_points(mask_u8: np.ndarray, use_cv2: bool = True):
    """data_collector.py:461-487: points of the largest outer contour where the turn angle
    |atan2(v1 x v2, v1 . v2)| is below pi/4.  Border steps are the eight unit moves, whose mutual angles
    are multiples of pi/4, and arctan2(1, 1) is not below np.pi / 4: the test holds exactly where the
    border doubles back (previous point == next point), or the contour is a single pixel."""
    mask_u8 = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    contours, _ = cv2.findContours(mask_u8, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)
    if not contours:
        return []
    contour = max(contours, key=cv2.contourArea)
    if use_cv2:
        pts = [tuple(p) for p in contour.reshape(-1, 2).tolist()]
    else:
        pts = outer_border(mask_u8, int(contour[0, 0, 0]), int(contour[0, 0, 1]))
    n = len(pts)
    return [pts[i] for i in range(n) if pts[i - 1] == pts[(i + 1) % n]]


def collector_extract(x, y, mask_u8, depth, scores):
    """data_collector.py:91-173 (_extract_patches): raw 32x32 windows by plain slicing, or None.
    Python slicing decides what "out of bounds" means: a negative start wraps and gives an empty or
    short window, an end beyond the image a short one; both fail the shape test."""
    h = PATCH // 2
    depth = np.asarray(depth, dtype=np.float32)
    dp = depth[y - h:y + h, x - h:x + h]
    mp = np.asarray(mask_u8)[y - h:y + h, x - h:x + h] != 0
    if dp.size == 0 or dp.shape != (PATCH, PATCH) or mp.shape != (PATCH, PATCH):
        return None
    if not np.isfinite(dp).all() or not mp.any():
        return None
    sp = []
    for name in SCORE_CHANNELS:
        p = np.asarray(scores[name])[y - h:y + h, x - h:x + h]
        if not np.isfinite(p).all():
            return None
        sp.append(p.astype(np.float32))
    return dp.copy(), mp.astype(np.float32), np.stack(sp)


def collector_rotate_point(point, angle, size=PATCH):
    """data_collector.py:402-418 (the point is rotated about (16,16) in whatever frame it is given in)."""
    x, y = point
    c = size // 2
    a = np.radians(angle)
    x -= c
    y -= c
    nx = x * np.cos(a) - y * np.sin(a)
    ny = x * np.sin(a) + y * np.cos(a)
    return (int(nx + c), int(ny + c))


def collect_sample(mask_u8, depth, scores, grasp_point, total_score, rng: CollectorRng, use_cv2=True, trace=None):
    """data_collector.py:175-348 without the bookkeeping: the samples one call appends, in order.
    Each sample is a dict(kind, label, is_augmented, grasp_point, total_score, patch float32 [9,32,32])
    with channels depth, mask, then SCORE_CHANNELS - the layout dataset.py stacks for training.
    Returns None where the reference returns False before adding anything.  `trace` (a list) receives one
    (attempt, kind, point, accepted) per negative pick."""
    mask_u8 = np.ascontiguousarray(mask_u8, dtype=np.uint8)
    H, W = mask_u8.shape
    x, y = int(grasp_point[0]), int(grasp_point[1])
    h = PATCH // 2
    if x < 0 or y < 0 or x >= W or y >= H:
        return None
    if y < h or y >= H - h or x < h or x >= W - h:               # _check_boundaries, :83-89
        return None
    got = collector_extract(x, y, mask_u8, depth, scores)
    if got is None:
        return None
    dp, mp, sp = got
    out = [dict(kind=KIND_POSITIVE, label=1, is_augmented=False, grasp_point=(x, y),
                total_score=float(total_score), patch=np.concatenate([dp[None], mp[None], sp]))]
    for k in (1, 2, 3):                                          # :250-299
        rd = np.rot90(dp, k)
        rm = (np.rot90(mp, k) > 0.5).astype(np.float32)
        rs = np.rot90(sp, k, axes=(-2, -1))
        factor = np.float32(rng.noise_factor(k))
        mean = torch.from_numpy(np.ascontiguousarray(rd)).mean().numpy()          # torch's float32 mean
        noisy = np.maximum(rd + rng.normal_patch(k) * np.float32(factor * mean), np.float32(0))
        out.append(dict(kind=KIND_POSITIVE + k, label=1, is_augmented=True,
                        grasp_point=collector_rotate_point((x, y), 90 * k),
                        total_score=float(total_score * rng.score_jitter(k)),
                        patch=np.concatenate([noisy[None], rm[None], rs]).astype(np.float32)))
    sets = (collector_tip_points(mask_u8, use_cv2), collector_stem_points(mask_u8, use_cv2),
            collector_edge_points(mask_u8, use_cv2))
    collected = 0
    for attempt in range(COLLECTOR_ATTEMPTS):                    # :309-341
        if collected >= COLLECTOR_NEG_MAX:
            break
        for kind, pts in enumerate(sets):
            if not pts or collected >= COLLECTOR_NEG_MAX:
                continue
            px, py = pts[rng.pick(attempt, kind, len(pts))]
            got = collector_extract(px, py, mask_u8, depth, scores)
            if trace is not None:
                trace.append((attempt, kind, (int(px), int(py)), got is not None))
            if got is None:
                continue
            dpn, mpn, spn = got
            out.append(dict(kind=KIND_TIP + kind, label=0, is_augmented=False, grasp_point=(int(px), int(py)),
                            total_score=0.0, patch=np.concatenate([dpn[None], mpn[None], spn])))
            collected += 1
    return out
