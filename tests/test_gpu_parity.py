"""GPU parity tests (run with ``-m gpu`` on the B200 box): every stage of the CUDA path, called through
the C-ABI, against oracle/leafgrasp_oracle.py on the same seeded inputs, and against the golden vectors
the reference itself produced (tests/golden/).  Nothing here reads /root/reference.

Bars (BASELINE.json north_star): distance transforms, masks and candidate indices bit-exact; score maps
within 1e-5 relative (they are in fact held to a few ulp); CNN logits within 1e-4 in fp32 / 1e-2 in bf16.
"""
import json
import os

import cv2
import numpy as np
import pytest
import torch

import leafgrasp_oracle as O
from leafgrasp_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLD, "golden_meta.json")))
SEED = META["config_seed"]
F64_MAPS = ("sdf_score", "approach_score", "isolation_map", "accessibility_map", "traditional_score")


def _engine(frames, H, W, labels=128):
    from leafgrasp_b200 import GraspEngine
    return GraspEngine(frames, H, W, labels)


def _cam(spec):
    from leafgrasp_b200 import camera_from_projection
    return camera_from_projection(synth.projection_matrix(spec))


def _blobs(rng, H, W, n):
    m = np.zeros((H, W), np.uint8)
    for _ in range(n):
        c = (int(rng.integers(0, W)), int(rng.integers(0, H)))
        ax = (int(rng.integers(2, max(3, W // 3))), int(rng.integers(2, max(3, H // 3))))
        cv2.ellipse(m, c, ax, float(rng.uniform(0, 180)), 0, 360, 1, -1)
    return m


@pytest.fixture(scope="module")
def state_dict():
    return O.seeded_state_dict(META["cnn_seed"])


@pytest.fixture(scope="module")
def blob(state_dict):
    from leafgrasp_b200 import pack_weights
    return pack_weights(state_dict)


# ------------------------------------------------------------------------------------------------------
# distance transforms
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(8, 8), (9, 700), (40, 57), (97, 131), (360, 480), (1080, 1440)])
def test_chamfer_bit_exact(shape):
    H, W = shape
    rng = np.random.default_rng(H * 7 + W)
    masks = [_blobs(rng, H, W, k + 1) for k in range(3)]
    masks.append(np.ones((H, W), np.uint8))                  # no source at all: OpenCV's saturation value
    m1 = np.ones((H, W), np.uint8)
    m1[H // 2, W // 3] = 0
    masks.append(m1)                                         # single far source
    m = np.stack(masks)
    eng = _engine(len(masks), H, W, 2)
    for invert in (False, True):
        dist, q16, mx = eng.chamfer(torch.from_numpy(m), invert=invert)
        dist, q16, mx = dist.cpu().numpy(), q16.cpu().numpy().view(np.uint32), mx.cpu().numpy().view(np.uint32)
        for k in range(len(masks)):
            src = (1 - m[k]) if invert else m[k]
            ref = cv2.distanceTransform(np.ascontiguousarray(src), cv2.DIST_L2, 5)
            np.testing.assert_array_equal(dist[k], ref, err_msg=f"mask {k} invert {invert}")
            if H * W <= 200 * 300:
                np.testing.assert_array_equal(q16[k], O.chamfer5_q16(src))
            assert mx[k] == q16[k].max()
    eng.close()


@pytest.mark.parametrize("shape", [(17, 23), (64, 96), (360, 480), (1080, 1440)])
def test_edt_squared_exact(shape):
    H, W = shape
    rng = np.random.default_rng(H + W)
    masks = np.stack([1 - _blobs(rng, H, W, k + 1) for k in range(3)])   # zeros (sources) = blobs
    eng = _engine(3, H, W, 2)
    d2, am = eng.edt_squared(torch.from_numpy(masks))
    d2, am = d2.cpu().numpy().view(np.uint32), am.cpu().numpy()
    for k in range(3):
        ref = O.edt_squared(masks[k])
        if (masks[k] == 0).any():
            np.testing.assert_array_equal(d2[k].astype(np.int64), ref)
            assert am[k] == int(ref.argmax())
        if H * W <= 64 * 96:
            np.testing.assert_array_equal(ref, O.edt_squared_bruteforce(masks[k]))
    eng.close()


@pytest.mark.parametrize("shape,frames", [((70, 90), 5), ((360, 480), 12), ((1080, 1440), 40)])
def test_edt_argmax_pruned_search(shape, frames):
    """The arg-max-only path (rows pruned against the running frame maximum) returns the first maximum of the
    exact field, for any scheduling: checked on batches large enough that every CTA walks several rows."""
    H, W = shape
    rng = np.random.default_rng(3 * H + W)
    masks = np.stack([1 - _blobs(rng, H, W, 1 + k % 7) for k in range(frames)])
    masks[0] = 1
    masks[0, H // 3, W // 5] = 0          # a single source pixel: the maximum sits in a corner
    eng = _engine(frames, H, W, 2)
    for _ in range(2):
        _, am = eng.edt_squared(torch.from_numpy(masks), argmax_only=True)
        am = am.cpu().numpy()
        for k in range(frames):
            if (masks[k] == 0).any():
                assert am[k] == int(O.edt_squared(masks[k]).argmax()), f"frame {k}"
    eng.close()


def test_outside_maximum_branch_and_bound():
    """sdf normalisation = max(inside, outside) chamfer distance (grasp_point_selector.py:531-532).  The outside
    maximum comes from the branch-and-bound search (or from the sweeps when its lists overflow): checked against the
    oracle's two-sweep transform on shapes chosen to break geometric shortcuts (rings, several components, thin
    lines, a mask covering the frame, a comb with thousands of boundary pixels -> fallback)."""
    H, W = 360, 480
    rng = np.random.default_rng(21)
    masks = []
    for k in range(10):
        masks.append(_blobs(rng, H, W, 1 + k % 4))
    ring = np.zeros((H, W), np.uint8); cv2.circle(ring, (240, 180), 150, 1, 9); masks.append(ring)
    two = np.zeros((H, W), np.uint8); cv2.circle(two, (30, 40), 25, 1, -1); cv2.circle(two, (450, 330), 22, 1, -1); masks.append(two)
    lines = np.zeros((H, W), np.uint8); cv2.line(lines, (5, 350), (470, 8), 1, 2); cv2.line(lines, (0, 0), (100, 300), 1, 1); masks.append(lines)
    masks.append(np.ones((H, W), np.uint8))
    corner = np.zeros((H, W), np.uint8); corner[:40, :60] = 1; masks.append(corner)
    comb = np.zeros((H, W), np.uint8); comb[40:300:2, 50:430] = 1; masks.append(comb)          # ~70k boundary pixels
    masks = np.stack(masks)
    dep = np.full(masks.shape, 0.5, np.float32)
    eng = _engine(len(masks), H, W, 2)
    res = eng.select_grasp_point(torch.from_numpy(masks), torch.from_numpy(dep), _cam(synth.SMALL))
    for k, m in enumerate(masks):
        di = O.chamfer5_q16(m).max()
        do = O.chamfer5_q16(1 - m).max() if (m == 0).any() else 0
        want = np.float32(max(int(di), int(do))) * np.float32(1.0 / 65536.0)
        assert np.float32(res[k]["sdf_max"]) == want, f"mask {k}: {res[k]['sdf_max']} vs {want}"
    eng.close()


# ------------------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec_name,count", [("SMALL", 4), ("CFG1", 1), ("CFG2", 2)])
def test_select_leaf_against_oracle(spec_name, count):
    spec = getattr(synth, spec_name)
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_batch(spec, SEED, 0, count)
    eng = _engine(count, spec.height, spec.width, 128)
    ids, rec = eng.select_leaf(torch.from_numpy(lab), torch.from_numpy(dep), _cam(spec))
    for i in range(count):
        o = O.select_optimal_leaf(lab[i], dep[i], P[0, 0], P[0, 2], P[1, 2])
        assert ids[i] == (o["leaf_id"] if o["leaf_id"] is not None else -1)
        present = rec[i][rec[i]["area"] > 0]
        assert [int(r["leaf_id"]) for r in present if r["is_tall"]] == o["tall"]
        for r in present:
            assert np.float32(r["median_depth"]) == np.float32(o["medians"][int(r["leaf_id"])])   # bit-exact median
        by_id = {c["leaf_id"]: c for c in o["candidates"]}
        for r in present:
            if not r["is_candidate"]:
                assert int(r["leaf_id"]) not in by_id
                continue
            c = by_id[int(r["leaf_id"])]
            assert r["area"] == c["area"]
            assert r["centroid_x"] == c["centroid"][0] and r["centroid_y"] == c["centroid"][1]
            np.testing.assert_allclose([r["clutter"], r["distance"], r["visibility"]], c["scores"], rtol=1e-6, atol=1e-12)
            np.testing.assert_allclose(r["mean_depth"], c["mean_depth"], rtol=3e-7)
    eng.close()


def test_select_leaf_matches_reference_golden():
    for f in META["frames"]:
        spec = getattr(synth, f["spec"])
        lab, dep = synth.make_frame(spec, SEED, f["index"])
        g = np.load(os.path.join(GOLD, f["file"]))
        eng = _engine(1, spec.height, spec.width, 128)
        ids, rec = eng.select_leaf(torch.from_numpy(lab), torch.from_numpy(dep), _cam(spec))
        assert ids[0] == int(g["leaf_id"])
        tall = [int(r["leaf_id"]) for r in rec[0] if r["area"] > 0 and r["is_tall"]]
        assert tall == g["tall"].tolist()
        eng.close()


# ------------------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------------------
def _ulp_diff32(a, b):
    ai = a.astype(np.float32).view(np.int32).astype(np.int64)
    bi = b.astype(np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


@pytest.mark.parametrize("spec_name,idx", [("SMALL", 0), ("SMALL", 1), ("SMALL", 3), ("CFG2", 0)])
def test_score_maps_against_strict_oracle(spec_name, idx):
    spec = getattr(synth, spec_name)
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_frame(spec, SEED, idx)
    leaf = O.select_optimal_leaf(lab, dep, P[0, 0], P[0, 2], P[1, 2])["leaf_id"]
    mask = (lab == leaf).astype(np.uint8)
    eng = _engine(1, spec.height, spec.width, 2)
    got = eng.score_maps(torch.from_numpy(mask), torch.from_numpy(dep), _cam(spec))
    ref = O.score_maps(mask, dep, P[0, 0], P[0, 2], P[1, 2], "strict")
    # integer / mask results: bit-exact
    np.testing.assert_array_equal(got["distance_map"][0].cpu().numpy(), ref["distance_map"])
    np.testing.assert_array_equal(got["stem_penalty"][0].cpu().numpy(), ref["stem_penalty"])
    np.testing.assert_array_equal(got["valid"][0].cpu().numpy().astype(bool), O.valid_regions(mask, ref))
    # the restated orientation: same float32 arithmetic on both sides
    assert abs(float(got["angle"][0]) - ref["_parts"]["angle"]) <= 1e-15
    # flatness: float32, every op correctly rounded except exp -> at most 1 ulp, and rarely
    fl = got["flatness_map"][0].cpu().numpy()
    ulp = _ulp_diff32(fl, ref["flatness_map"])
    assert ulp.max() <= 1 and (ulp > 0).mean() < 1e-3
    for k in F64_MAPS:
        np.testing.assert_allclose(got[k][0].cpu().numpy(), ref[k], rtol=1e-9, atol=1e-12, err_msg=k)
    # and within the north-star tolerance of what the reference itself produced (golden samples of its maps; the
    # oracle's "reference" arithmetic is not used here: torch's CPU convolution picks its algorithm per host CPU)
    fn = [f["file"] for f in META["frames"] if f["spec"] == spec_name and f["index"] == idx][0]
    g = np.load(os.path.join(GOLD, fn))
    assert int(g["leaf_id"]) == leaf
    yx = g["sample_yx"]
    for k in F64_MAPS + ("flatness_map",):
        np.testing.assert_allclose(got[k][0].cpu().numpy()[yx[:, 0], yx[:, 1]].astype(np.float64), g["sample_" + k],
                                   rtol=1e-5, atol=1e-6, err_msg=k)
    eng.close()


def test_candidate_points_exact():
    rng = np.random.default_rng(3)
    H, W = 240, 320
    eng = _engine(3, H, W, 2)
    score = rng.random((3, H, W))
    valid = np.ones((3, H, W), np.uint8)
    valid[1] = 0
    valid[1, 100:140, 100:150] = 1          # few positives -> zero-key fill
    valid[2, :, :] = 0                      # nothing valid at all
    xy, cnt = eng.candidate_points(torch.from_numpy(score), torch.from_numpy(valid))
    xy, cnt = xy.cpu().numpy(), cnt.cpu().numpy()
    for k in range(3):
        ref = O.candidate_points(score[k], valid[k].astype(bool))
        assert cnt[k] == len(ref)
        assert [tuple(p) for p in xy[k, :cnt[k]].tolist()] == ref
    eng.close()


def test_orientation_against_opencv():
    rng = np.random.default_rng(0)
    H, W = 300, 400
    masks = []
    for t in range(24):
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(1, 4))):
            cv2.ellipse(m, (int(rng.integers(60, 340)), int(rng.integers(50, 250))),
                        (int(rng.integers(15, 90)), int(rng.integers(8, 60))), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        if t % 3 == 0:   # occluder splits the leaf into components
            cv2.ellipse(m, (int(rng.integers(60, 340)), int(rng.integers(50, 250))), (120, 6), float(rng.uniform(0, 180)), 0, 360, 0, -1)
        masks.append(m)
    masks = np.stack(masks)
    eng = _engine(len(masks), H, W, 2)
    out = eng.leaf_orientation(torch.from_numpy(masks)).cpu().numpy()
    for k, m in enumerate(masks):
        ang_cv = O.leaf_orientation(m)
        ang_re = O.leaf_orientation_restated(m)
        assert out[k, 0] == ang_re[0], (k, out[k, 0], ang_re[0])                 # same restated arithmetic
        assert abs(out[k, 0] - ang_cv[0]) < 1e-6                                 # OpenCV within float32 noise
        np.testing.assert_allclose(out[k, 1:3], [ang_re[1], ang_re[2]], rtol=1e-6)
    eng.close()


def test_orientation_largest_contour_rules():
    """Which component wins (cv2: max contourArea over the external contours, first maximum in cv2's list order), on
    masks where pixel count and contour area disagree: the kernel skips the border trace when a lower bound of the
    biggest component's contour area already exceeds every other component's bounding box, and traces otherwise."""
    H, W = 240, 320
    masks = []
    m = np.zeros((H, W), np.uint8)                       # ring (few pixels, large contour) vs solid disc (more pixels)
    cv2.circle(m, (90, 120), 70, 1, 6); cv2.circle(m, (240, 120), 40, 1, -1); masks.append(m)
    m = np.zeros((H, W), np.uint8)                       # two identical rectangles: equal areas, the later start wins
    m[30:90, 40:140] = 1; m[130:190, 170:270] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8)                       # the same with the first one a pixel larger
    m[30:91, 40:140] = 1; m[130:190, 170:270] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8)                       # one big leaf and specks: no trace needed
    cv2.ellipse(m, (160, 120), (120, 60), 25.0, 0, 360, 1, -1)
    for (x, y) in ((5, 5), (300, 20), (20, 220), (310, 230), (160, 5)):
        m[y:y + 2, x:x + 3] = 1
    masks.append(m)
    m = np.zeros((H, W), np.uint8)                       # thin diagonal line (area 0) and a small blob
    for t in range(150):
        m[20 + t, 30 + t] = 1
    m[200:206, 250:258] = 1; masks.append(m)
    m = np.zeros((H, W), np.uint8); m[100, 100] = 1; masks.append(m)            # a single pixel
    m = np.zeros((H, W), np.uint8); m[0:H, 0:50] = 1; m[0:40, 0:W] = 1; masks.append(m)   # L-shape along the image border
    m = np.zeros((H, W), np.uint8)                       # comb: many pixels on the border, tiny interior
    m[60:180:2, 40:280] = 1; m[60:180, 40:42] = 1; cv2.circle(m, (160, 215), 18, 1, -1); masks.append(m)
    m = np.zeros((H, W), np.uint8)                       # two rings, nested bounding boxes
    cv2.circle(m, (160, 120), 100, 1, 3); cv2.circle(m, (160, 120), 60, 1, 3); masks.append(m)
    masks = np.stack(masks)
    eng = _engine(len(masks), H, W, 2)
    out = eng.leaf_orientation(torch.from_numpy(masks)).cpu().numpy()
    for k, m in enumerate(masks):
        ang_re = O.leaf_orientation_restated(m)
        assert out[k, 0] == ang_re[0], (k, out[k, 0], ang_re[0])
        np.testing.assert_allclose(out[k, 1:3], [ang_re[1], ang_re[2]], rtol=1e-6, err_msg=f"mask {k}")
        np.testing.assert_allclose(out[k, 3:5], ang_re[3], rtol=1e-6, atol=1e-4, err_msg=f"mask {k}")
    eng.close()


# ------------------------------------------------------------------------------------------------------
# CNN
# ------------------------------------------------------------------------------------------------------
def test_cnn_fp32_against_reference_logits(blob):
    g = np.load(os.path.join(GOLD, "cnn_patches.npz"))
    eng = _engine(1, 64, 64, 2)
    eng.set_cnn_weights(blob)
    y = eng.cnn_forward(torch.from_numpy(g["x"])).cpu().numpy()
    np.testing.assert_allclose(y, g["logits"], atol=1e-4, rtol=1e-4)
    eng.close()


def test_cnn_dropin_module(state_dict):
    from leafgrasp_b200 import GraspPointCNN
    g = np.load(os.path.join(GOLD, "cnn_patches.npz"))
    net = GraspPointCNN(in_channels=9)
    net.load_state_dict(state_dict)
    net.eval()
    with torch.no_grad():
        y = net(torch.from_numpy(g["x"]).cuda())
    assert y.shape == (16, 1)
    np.testing.assert_allclose(y.reshape(-1).cpu().numpy(), g["logits"], atol=1e-4, rtol=1e-4)


def _torch_layer_features(sd, x, layer):
    """fp32 torch activations after conv `layer` (BN + ReLU applied; pooled after layers 1, 3, 5)."""
    import torch.nn.functional as F
    k = 0
    for b in range(3):
        for conv, bn in ((0, 1), (3, 4)):
            x = F.conv2d(x, sd[f"encoder.{b}.{conv}.weight"], sd[f"encoder.{b}.{conv}.bias"], padding=1)
            x = F.batch_norm(x, sd[f"encoder.{b}.{bn}.running_mean"], sd[f"encoder.{b}.{bn}.running_var"],
                             sd[f"encoder.{b}.{bn}.weight"], sd[f"encoder.{b}.{bn}.bias"], False, 0.0, 1e-5)
            x = F.relu(x)
            if conv == 3:
                x = F.max_pool2d(x, 2)
            if k == layer:
                return x
            k += 1
    raise ValueError(layer)


@pytest.mark.parametrize("layer", [0, 1, 2, 3, 4, 5])
def test_cnn_bf16_layer_features(layer, state_dict, blob):
    """Every tcgen05 conv layer against torch fp32 on the same patches: bf16 operands, fp32 accumulation."""
    g = np.load(os.path.join(GOLD, "cnn_patches.npz"))
    x = torch.from_numpy(g["x"])
    rng = np.random.default_rng(5)
    extra = torch.from_numpy(rng.random((45, 9, 32, 32), dtype=np.float32))   # 61 patches: several work items
    extra[:, 1] = (extra[:, 1] > 0.5).float()
    x = torch.cat([x, extra])
    eng = _engine(1, 64, 64, 2)
    eng.set_cnn_weights(blob)
    got = eng.cnn_bf16_features(x, layer).cpu()
    with torch.no_grad():
        want = _torch_layer_features(state_dict, x, layer)
    if layer == 5:
        want = want.permute(0, 2, 3, 1).contiguous()
    assert got.shape == want.shape
    scale = float(want.abs().max())
    err = float((got - want).abs().max())
    assert err <= 2e-2 * scale, f"layer {layer}: max abs err {err} vs activation scale {scale}"
    eng.close()


def test_cnn_bf16_against_reference_logits(blob, state_dict):
    """north_star: CNN logits within 1e-2 in bf16 (against the reference's fp32 logits, golden vectors)."""
    g = np.load(os.path.join(GOLD, "cnn_patches.npz"))
    eng = _engine(1, 64, 64, 2)
    eng.set_cnn_weights(blob)
    y = eng.cnn_forward(torch.from_numpy(g["x"]), use_bf16=True).cpu().numpy()
    np.testing.assert_allclose(y, g["logits"], atol=1e-2, rtol=1e-2)
    # a batch that spans several activation chunks agrees with the fp32 CUDA path
    rng = np.random.default_rng(9)
    x = torch.from_numpy(rng.random((2300, 9, 32, 32), dtype=np.float32))
    x[:, 1] = (x[:, 1] > 0.5).float()
    y16 = eng.cnn_forward(x, use_bf16=True).cpu().numpy()
    y32 = eng.cnn_forward(x, use_bf16=False).cpu().numpy()
    np.testing.assert_allclose(y16, y32, atol=1e-2, rtol=1e-2)
    eng.close()


def test_cnn_architecture_variants_against_reference_logits():
    """every architecture of the reference's sweep through the drop-in class on the GPU (fp32
    kernels), against the logits the reference's own GraspPointCNN produced for the same weights (golden)."""
    from leafgrasp_b200 import GraspPointCNN
    meta = json.load(open(os.path.join(GOLD, "cnn_variants.json")))
    gold = np.load(os.path.join(GOLD, "cnn_variants.npz"))
    x = torch.from_numpy(gold["x"]).cuda()
    for i, v in enumerate(meta["variants"]):
        net = GraspPointCNN(in_channels=9, attention_type=v["attention_type"], encoder_filters=v["encoder_filters"])
        net.load_state_dict(O.seeded_state_dict_from_shapes(v["shapes"], v["seed"]))
        net.eval()
        net.use_bf16 = False
        with torch.no_grad():
            y = net(x).reshape(-1).cpu().numpy()
        np.testing.assert_allclose(y, gold[f"logits_{i}"], atol=2e-4, rtol=2e-4, err_msg=str((v["attention_type"], v["encoder_filters"])))
        # the attention variants of the encoder [64, 128, 256] also run their convolutions on the tensor cores (bf16);
        # for the other encoders the switch is ignored and the fp32 kernels run
        net.use_bf16 = True
        with torch.no_grad():
            y16 = net(x).reshape(-1).cpu().numpy()
        if v["encoder_filters"] == [64, 128, 256]:
            assert not np.array_equal(y16, y)
            np.testing.assert_allclose(y16, gold[f"logits_{i}"], atol=1e-2, err_msg="bf16 " + str(v["attention_type"]))
        else:
            np.testing.assert_array_equal(y16, y)
        net.use_bf16 = False
        # a batch that does not fit one activation chunk gives the same rows
        with torch.no_grad():
            yb = net(x.repeat(500, 1, 1, 1)).reshape(500, -1).cpu().numpy()
        np.testing.assert_allclose(yb, np.tile(y, (500, 1)), atol=1e-6)


def test_cfg4_cnn_only_65536_patches(blob):
    """BASELINE config[3]: 65 536 patches through the tensor-core path (13 activation chunks of 5 120) against the
    fp32 CUDA-core path, logits within the bf16 bar."""
    eng = _engine(256, 64, 64, 2)
    eng.set_cnn_weights(blob)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((65536, 9, 32, 32), generator=g, device="cuda")
    x[:, 1] = (x[:, 1] > 0.5).float()
    y16 = eng.cnn_forward(x, use_bf16=True)
    y32 = eng.cnn_forward(x, use_bf16=False)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); eng.cnn_forward(x, use_bf16=True); e1.record(); eng.cnn_forward(x, use_bf16=False); e2.record()
    torch.cuda.synchronize()
    t16, t32 = e0.elapsed_time(e1), e1.elapsed_time(e2)
    print(f"cfg4: 65536 patches bf16 tcgen05 {t16:.1f} ms ({65536 * 312.83e6 / t16 / 1e9:.0f} TFLOP/s), fp32 CUDA cores {t32:.1f} ms")
    np.testing.assert_allclose(y16.cpu().numpy(), y32.cpu().numpy(), atol=1e-2, rtol=1e-2)
    assert t16 < t32
    eng.close()


# ------------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------------
def _check_frame(r, lab, dep, P, sd, patches=None):
    o = O.process_frame(lab, dep, P, sd, arith="strict")
    assert r["leaf_id"] == (o["leaf_id"] if o["leaf_id"] is not None else -1)
    if o["leaf_id"] is None:
        return o
    d = o["debug"]
    n = len(d["picks"])
    assert r["n_candidates"] == n
    got = list(zip(r["cand_x"][:n].tolist(), r["cand_y"][:n].tolist()))
    assert got == d["picks"], "candidate pixels differ from the strict oracle"
    np.testing.assert_allclose(r["trad"][:n], d["trad_at"], rtol=1e-9)
    for k in range(n):
        has = d["logits"][k] is not None
        assert bool(r["ml_valid"][k]) == has
        if has:
            assert abs(r["logit"][k] - d["logits"][k]) < 2e-4
            assert abs(r["ml"][k] - d["ml"][k]) < 1e-4
            if patches is not None:
                ref_patch = O.patch_tensor((lab == o["leaf_id"]).astype(np.uint8), dep, d["scores"], *d["picks"][k])
                np.testing.assert_allclose(patches[k], ref_patch, rtol=1e-5, atol=2e-6)
    best, p3, pre = o["grasp"]
    assert (int(r["grasp_x"]), int(r["grasp_y"])) == tuple(best)
    np.testing.assert_allclose(r["grasp_3d"], np.array(p3, dtype=np.float64), rtol=1e-12)
    np.testing.assert_allclose(r["pre_grasp"], np.array(pre, dtype=np.float64), rtol=1e-9)
    return o


@pytest.mark.parametrize("spec_name,count", [("SMALL", 4), ("CFG1", 1), ("CFG2", 3)])
def test_process_batch_against_strict_oracle(spec_name, count, state_dict, blob):
    spec = getattr(synth, spec_name)
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_batch(spec, SEED, 0, count)
    eng = _engine(count, spec.height, spec.width, 128)
    eng.set_cnn_weights(blob)
    res = eng.process_batch(torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    patches = eng.last_patches(count).cpu().numpy()
    for i in range(count):
        _check_frame(res[i], lab[i], dep[i], P, state_dict, patches[i])
    # host-buffer entry point gives the same records
    res_h = eng.process_batch_host(torch.from_numpy(lab).pin_memory(), torch.from_numpy(dep).pin_memory(), _cam(spec))
    for name in ("leaf_id", "n_candidates", "cand_x", "cand_y", "grasp_x", "grasp_y"):
        np.testing.assert_array_equal(res[name], res_h[name])
    eng.close()


def test_process_batch_against_reference_golden(blob):
    """The reference's own outputs: every candidate whose key is positive, the fused pick and the 3-D points."""
    for f in META["frames"]:
        spec = getattr(synth, f["spec"])
        lab, dep = synth.make_frame(spec, SEED, f["index"])
        g = np.load(os.path.join(GOLD, f["file"]))
        eng = _engine(1, spec.height, spec.width, 128)
        eng.set_cnn_weights(blob)
        r = eng.process_batch(torch.from_numpy(lab)[None].cuda(), torch.from_numpy(dep)[None].cuda(), _cam(spec))[0]
        assert r["leaf_id"] == int(g["leaf_id"])
        n_pos = int(g["n_positive"])
        assert r["n_positive"] == n_pos
        got = np.stack([r["cand_x"][:n_pos], r["cand_y"][:n_pos]], axis=1)
        np.testing.assert_array_equal(got, g["candidates"][:n_pos])
        np.testing.assert_allclose(r["trad"][:n_pos], g["trad_at"][:n_pos], rtol=1e-5)
        assert abs(r["angle"] - float(g["angle"])) < 1e-6
        assert (int(r["grasp_x"]), int(r["grasp_y"])) == tuple(g["grasp_2d"].tolist())
        np.testing.assert_allclose(r["grasp_3d"], g["grasp_3d"], rtol=1e-12)
        np.testing.assert_allclose(r["pre_grasp"], g["pre_grasp"], rtol=1e-9)
        eng.close()


def test_cfg3_4k_dense_clutter_frame(state_dict, blob):
    """BASELINE config[2]: one 3840x2160 frame with 100 overlapping leaves against the strict oracle."""
    spec = synth.CFG3
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_batch(spec, SEED, 0, 1)
    eng = _engine(1, spec.height, spec.width, 128)
    eng.set_cnn_weights(blob)
    res = eng.process_batch(torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    _check_frame(res[0], lab[0], dep[0], P, state_dict)
    eng.close()


_EXACT_FIELDS = ("status", "leaf_id", "n_candidates", "n_positive", "cand_x", "cand_y", "trad", "ml_valid", "best_index",
                 "ml_used", "grasp_x", "grasp_y", "grasp_3d", "pre_grasp", "angle", "sdf_max", "region")


def test_batch_invariance_and_scheduling_independence(blob):
    """Size-independent properties at the benchmark's frame size: a frame's record does not depend on the batch it
    travels in, on its position in the batch, on stage overlap (side stream on/off), or on the run (atomics and
    pruning make the schedule vary; results may not).  fp32 CNN, so logits are bit-identical too."""
    spec = synth.CFG2
    n = 24
    lab, dep = synth.make_batch(spec, SEED, 100, 6)
    idx = np.arange(n) % 6
    labs, deps = torch.from_numpy(lab[idx]).cuda(), torch.from_numpy(dep[idx]).cuda()
    eng = _engine(n, spec.height, spec.width, 128)
    eng.set_cnn_weights(blob)
    a = eng.process_batch(labs, deps, _cam(spec))
    b = eng.process_batch(labs, deps, _cam(spec))                     # idempotence
    eng.set_overlap(False)
    c = eng.process_batch(labs, deps, _cam(spec))                     # serialised stages
    eng.set_overlap(True)
    single = eng.process_batch(labs[:1], deps[:1], _cam(spec))        # batch of one
    from leafgrasp_b200 import GraspEngine
    eng3 = GraspEngine(n, spec.height, spec.width, 128, lanes=3)      # three parts side by side on their own streams
    eng3.set_cnn_weights(blob)
    assert eng3.lane_split(n) == [(0, 8), (8, 16), (16, 24)] and eng3.lane_split(5) == [(0, 5)]
    d = eng3.process_batch(labs, deps, _cam(spec))
    eng3.close()
    for name in _EXACT_FIELDS + ("logit",):
        np.testing.assert_array_equal(a[name], b[name], err_msg=name)
        np.testing.assert_array_equal(a[name], c[name], err_msg=name)
        np.testing.assert_array_equal(a[name], d[name], err_msg=name)
        np.testing.assert_array_equal(a[name][:1], single[name], err_msg=name)
        for k in range(6, n):                                          # same frame, other slot
            np.testing.assert_array_equal(a[name][k], a[name][k % 6], err_msg=name)
    assert (a["leaf_id"] > 0).all() and (a["n_candidates"] == 20).all()
    eng.close()


@pytest.mark.parametrize("shape", [(203, 317), (360, 484), (97, 1000)])
def test_odd_sizes_take_the_generic_kernels(shape, state_dict, blob):
    """Widths that are not multiples of 8 (no vector loads, generic chamfer kernel) against the strict oracle."""
    H, W = shape
    spec = synth.FrameSpec(height=H, width=W, n_leaves=3, a_range=(70.0, 90.0), b_range=(45.0, 60.0), margin=min(H, W) * 0.3)
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_batch(spec, SEED, 7, 2)
    eng = _engine(2, H, W, 16)
    eng.set_cnn_weights(blob)
    res = eng.process_batch(torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    for i in range(2):
        _check_frame(res[i], lab[i], dep[i], P, state_dict)
    eng.close()


def test_bf16_pipeline_matches_fp32_picks(blob):
    """The tensor-core CNN inside the whole path: same candidates, logits within the bf16 bar, and the fused pick
    agrees wherever the fp32 decision has a margin larger than the bf16 error."""
    spec = synth.CFG2
    lab, dep = synth.make_batch(spec, SEED, 200, 8)
    eng = _engine(8, spec.height, spec.width, 128)
    eng.set_cnn_weights(blob)
    labs, deps = torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda()
    r32 = eng.process_batch(labs, deps, _cam(spec), use_bf16=False)
    r16 = eng.process_batch(labs, deps, _cam(spec), use_bf16=True)
    for name in ("leaf_id", "n_candidates", "cand_x", "cand_y", "trad", "ml_valid"):
        np.testing.assert_array_equal(r32[name], r16[name], err_msg=name)
    ok = r32["ml_valid"] > 0
    np.testing.assert_allclose(r16["logit"][ok], r32["logit"][ok], atol=1e-2, rtol=1e-2)
    np.testing.assert_allclose(r16["ml"][ok], r32["ml"][ok], atol=5e-3)
    same = (r16["best_index"] == r32["best_index"]).mean()
    assert same >= 0.75, f"fused pick agreement {same}"
    eng.close()


def test_label_table_limits(state_dict, blob):
    """Label ids: a table of the maximum size (1024 ids, shared-memory table above the default limit) gives the same
    pick as a small one; an id outside the table is reported per frame (LG_ST_LABEL_RANGE) without disturbing the
    other frames of the batch."""
    spec = synth.SMALL
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_batch(spec, SEED, 0, 2)
    lab = lab.copy()
    lab[0][lab[0] == 2] = 1000                    # a legal id near the top of the table
    eng = _engine(2, spec.height, spec.width, 1024)
    eng.set_cnn_weights(blob)
    res = eng.process_batch(torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    for i in range(2):
        _check_frame(res[i], lab[i], dep[i], P, state_dict)
    eng.close()
    eng = _engine(2, spec.height, spec.width, 16)
    eng.set_cnn_weights(blob)
    bad = lab.copy()
    bad[0][bad[0] == 1000] = 999                  # outside a 16-entry table
    res = eng.process_batch(torch.from_numpy(bad).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    assert res[0]["status"] & 2 and res[0]["leaf_id"] == -1
    _check_frame(res[1], lab[1], dep[1], P, state_dict)
    eng.close()


def test_striped_leaf_overflows_the_column_run_table(state_dict, blob):
    """A leaf made of 60 thin horizontal stripes: every column crosses more leaf runs than leaf_stats_kernel records
    (STC_BND), so the distance-transform column pass takes its fallback.  Same records and pick as the oracle."""
    spec = synth.SMALL
    H, W = spec.height, spec.width
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_frame(spec, SEED, 1)
    lab = lab.copy()
    stripes = (np.arange(H) // 3) % 2 == 0
    lab[np.ix_(stripes, np.arange(40, 200))] = 7
    eng = _engine(1, H, W, 16)
    eng.set_cnn_weights(blob)
    ids, rec = eng.select_leaf(torch.from_numpy(lab)[None], torch.from_numpy(dep)[None], _cam(spec))
    o = O.select_optimal_leaf(lab, dep, P[0, 0], P[0, 2], P[1, 2])
    assert ids[0] == (o["leaf_id"] if o["leaf_id"] is not None else -1)
    by_id = {c["leaf_id"]: c for c in o["candidates"]}
    for r in rec[0][rec[0]["is_candidate"] > 0]:
        c = by_id[int(r["leaf_id"])]
        np.testing.assert_allclose([r["clutter"], r["distance"], r["visibility"]], c["scores"], rtol=1e-6, atol=1e-12)
    res = eng.process_batch(torch.from_numpy(lab)[None].cuda(), torch.from_numpy(dep)[None].cuda(), _cam(spec))
    _check_frame(res[0], lab, dep, P, state_dict)
    eng.close()


def test_empty_and_degenerate_frames(blob):
    spec = synth.SMALL
    H, W = spec.height, spec.width
    lab = np.zeros((3, H, W), np.int16)
    dep = np.full((3, H, W), 0.5, np.float32)
    lab[1, 10:60, 10:60] = 3          # one leaf, too small (2500 px < 10000)
    lab[2, :, :] = 5                  # a single id everywhere: it is the "background" (smallest id dropped)
    eng = _engine(3, H, W, 16)
    eng.set_cnn_weights(blob)
    res = eng.process_batch(torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda(), _cam(spec))
    P = synth.projection_matrix(spec)
    for i in range(3):
        o = O.select_optimal_leaf(lab[i], dep[i], P[0, 0], P[0, 2], P[1, 2])
        assert o["leaf_id"] is None
        assert res[i]["leaf_id"] == -1 and res[i]["n_candidates"] == 0
        assert res[i]["status"] & 1
    eng.close()


def test_dropin_classes_match_reference_golden(state_dict):
    """The call sequence of leaf_grasp_node_v3.py:110-119 through the drop-in classes."""
    from leafgrasp_b200 import GraspPointCNN, GraspPointSelector, ImageProcessor, OptimalLeafSelector
    spec = synth.SMALL
    P = synth.projection_matrix(spec)
    g = np.load(os.path.join(GOLD, "frame_small_2.npz"))
    lab, dep = synth.make_frame(spec, SEED, 2)
    dev = torch.device("cuda")
    scorer = OptimalLeafSelector(dev)
    scorer.set_camera_params(P)
    sel = GraspPointSelector(dev)
    sel.set_camera_params(P)
    net = GraspPointCNN(in_channels=9)
    net.load_state_dict(state_dict)
    net.eval()
    sel.ml_predictor = net
    ip = ImageProcessor(spec.height, spec.width, 21, 5)
    mask_t, depth_t = torch.from_numpy(lab).to(dev), torch.from_numpy(dep).to(dev)
    leaf = scorer.select_optimal_leaf(mask_t, depth_t)
    assert leaf == int(g["leaf_id"])
    assert scorer.get_tall_leaves() == g["tall"].tolist()
    g2, g3, pre = sel.select_grasp_point(mask_t == leaf, depth_t, ip)
    assert tuple(g2) == tuple(g["grasp_2d"].tolist())
    np.testing.assert_allclose(g3, g["grasp_3d"], rtol=1e-12)
    np.testing.assert_allclose(pre, g["pre_grasp"], rtol=1e-9)
    mask_np = (lab == leaf).astype(np.uint8)
    scores = sel._calculate_all_scores(mask_np, depth_t, ip)
    yx = g["sample_yx"]
    for k in ("sdf_score", "approach_score", "flatness_map", "isolation_map", "distance_map", "accessibility_map",
              "stem_penalty", "traditional_score"):
        np.testing.assert_allclose(scores[k][yx[:, 0], yx[:, 1]].astype(np.float64), g["sample_" + k], rtol=1e-5, atol=1e-6)
    valid = sel._get_valid_regions(mask_np, scores)
    assert int(valid.sum()) == int(g["valid_count"])
    cands = sel._get_candidate_points(scores["traditional_score"], valid, top_k=20, min_distance=10)
    n_pos = int(g["n_positive"])
    assert cands[:n_pos] == [tuple(p) for p in g["candidates"][:n_pos].tolist()]
    ang = sel.estimate_leaf_orientation(mask_np)
    assert abs(ang[0] - float(g["angle"])) < 1e-6
    ml = sel.get_ml_score(mask_t == leaf, depth_t, scores, cands[0])
    assert ml is not None and 0.5 <= ml <= 1.0
    # error convention: no camera -> (None, None, None), no exception
    bare = GraspPointSelector(dev)
    assert bare.select_grasp_point(mask_t == leaf, depth_t, ip) == (None, None, None)
