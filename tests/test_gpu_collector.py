"""GPU parity tests of the training-sample collector (csrc/lg_collect.cu), through the C-ABI:
against the golden vectors the reference's own EnhancedGraspDataCollector produced (tests/golden/collector.npz) and
against oracle.collect_sample on seeded masks.

Bars: candidate sets, picks, points, labels and the raw depth / mask / distance / isolation / stem windows bit-exact;
float score windows within 1e-5 (the score-map bar); noisy depth within 1e-6 absolute (torch's float32 mean and the
float64 Box-Muller normals are the only non-integer steps)."""
import cv2
import numpy as np
import pytest
import torch

import collector_util as CU
import leafgrasp_oracle as O
from leafgrasp_b200 import _native as N
from leafgrasp_b200 import synth

pytestmark = pytest.mark.gpu
EXACT_CH = (0, 1, 5, 6, 8)        # depth, mask, isolation ramp, chamfer distance, stem penalty
FLOAT_CH = (2, 3, 4, 7)           # sdf_score, approach_score, flatness_map, accessibility_map


def _engine(frames, H, W):
    from leafgrasp_b200 import GraspEngine
    return GraspEngine(frames, H, W, 128)


def _cam(P):
    from leafgrasp_b200 import camera_from_projection
    return camera_from_projection(P)


def _check_patch(got, want, augmented, tag):
    for ch in EXACT_CH:
        if ch == 0 and augmented:
            np.testing.assert_allclose(got[0], want[0], rtol=0, atol=1e-6, err_msg=f"{tag} noisy depth")
        else:
            assert np.array_equal(got[ch], want[ch]), f"{tag} channel {ch}"
    for ch in FLOAT_CH:
        np.testing.assert_allclose(got[ch], want[ch], rtol=1e-5, atol=1e-6, err_msg=f"{tag} channel {ch}")


@pytest.mark.parametrize("name", CU.CASES)
@pytest.mark.parametrize("defaults", [False, True])
def test_collector_against_reference_golden(name, defaults):
    mask, dep, P, g, lab, leaf = CU.case_inputs(name)
    H, W = mask.shape
    eng = _engine(1, H, W)
    d = torch.from_numpy(dep).cuda()
    if lab is not None:
        lt = torch.from_numpy(lab).cuda()
        res = eng.process_batch(lt, d, _cam(P))
        assert int(res["leaf_id"][0]) == leaf
        kw = dict(labels=lt)
    else:
        mt = torch.from_numpy(mask).cuda()
        res = eng.select_grasp_point(mt, d, _cam(P))
        kw = dict(mask=mt)
    assert (int(res["grasp_x"][0]), int(res["grasp_y"][0])) == tuple(g["grasp"].tolist())    # CV-only pick = candidate 0
    if not defaults:
        kw.update(grasp_xy=[g["grasp"].tolist()], total_score=[float(g["total_score"])])
    patches, meta, sizes = eng.collect_samples(d, seed=int(g["rng_seed"]), first_frame_index=int(g["frame_index"]), **kw)
    patches = patches.cpu().numpy()[0]
    meta = meta[0]
    assert sizes[0].tolist() == [len(g["set_tip"]), len(g["set_stem"]), len(g["set_edge"])]
    n = len(g["labels"])
    valid = [k for k in range(N.SAMPLES_PER_FRAME) if meta[k]["valid"]]
    assert len(valid) == n
    for i, k in enumerate(valid):
        m = meta[k]
        assert (int(m["label"]), int(m["is_augmented"])) == (int(g["labels"][i]), int(g["is_augmented"][i]))
        assert [int(m["x"]), int(m["y"])] == g["points"][i].tolist(), f"sample {i}"
        np.testing.assert_allclose(float(m["total_score"]), float(g["total_scores"][i]), rtol=1e-6 if defaults else 1e-15)
        _check_patch(patches[k], g["patches"][i], bool(m["is_augmented"]), f"{name} sample {i}")
    # the candidate sets themselves, in the reference's list order
    for kind, key in enumerate(("set_tip", "set_stem", "set_edge")):
        want = g[key]
        if len(want) == 0:
            continue
        got = eng.collector_points(kind, np.arange(len(want))[None], **{k: v for k, v in kw.items() if k in ("labels", "mask")})
        assert np.array_equal(got[0], want), key


def _adversarial_masks(H, W):
    """Masks that stress the three sets: plateaus (many equal local maxima), thin spikes and doubled-back borders,
    components touching the image border (windows of picks leave the image), stems reaching the last row."""
    rng = np.random.default_rng(77)
    out = []
    m = np.zeros((H, W), np.uint8); m[60:140, 40:220] = 1; out.append(m)                       # rectangle: ridge plateau
    m = np.zeros((H, W), np.uint8); m[120:H, 30:230] = 1; m[100:120, 128] = 1; out.append(m)   # touches bottom, spike
    m = np.zeros((H, W), np.uint8); cv2.ellipse(m, (130, 150), (110, 45), 20.0, 0, 360, 1, -1)
    m[150, 0:30] = 1; m[100:150:1, 200] = 1; out.append(m)                                      # ellipse + spikes to the border
    m = np.zeros((H, W), np.uint8); cv2.ellipse(m, (60, 170), (50, 25), 0.0, 0, 360, 1, -1)
    cv2.ellipse(m, (190, 60), (40, 30), 45.0, 0, 360, 1, -1); m[185, 110:150] = 1; out.append(m)   # two components + line
    for _ in range(3):
        m = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(2, 6))):
            c = (int(rng.integers(20, W - 20)), int(rng.integers(20, H - 20)))
            cv2.ellipse(m, c, (int(rng.integers(8, 80)), int(rng.integers(8, 50))), float(rng.uniform(0, 180)), 0, 360, 1, -1)
        for _ in range(4):                                                                     # random one-pixel whiskers
            x, y = int(rng.integers(5, W - 5)), int(rng.integers(5, H - 5))
            if rng.random() < 0.5:
                m[y, x:min(W, x + int(rng.integers(3, 25)))] = 1
            else:
                m[y:min(H, y + int(rng.integers(3, 25))), x] = 1
        out.append(m)
    m = np.zeros((H, W), np.uint8); m[170:H, 0:W] = 1; out.append(m)                          # full-width band at the bottom
    return np.stack(out)


def test_collector_sets_and_samples_vs_oracle():
    H, W = 200, 260
    masks = _adversarial_masks(H, W)
    n = len(masks)
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[:H, :W]
    depth = np.stack([(0.4 + 2e-4 * xx - 1e-4 * yy + rng.normal(0, 1e-3, (H, W))).astype(np.float32) for _ in range(n)])
    depth[masks == 0] = 0.8
    depth[0, 100:103, 150:153] = np.nan                      # a hole in the first leaf's depth: windows over it are rejected
    spec = synth.FrameSpec(height=H, width=W, n_leaves=1)
    P = synth.projection_matrix(spec)
    f, cx, cy = P[0, 0], P[0, 2], P[1, 2]
    eng = _engine(n, H, W)
    mt, dt = torch.from_numpy(masks).cuda(), torch.from_numpy(depth).cuda()
    eng.select_grasp_point(mt, dt, _cam(P))
    seed, first = 991, 40
    # explicit grasp points: the middle leaf pixel in raster order (frame 7's lies in column 0: no sample at all)
    grasp = []
    for b in range(n):
        ys, xs = np.nonzero(masks[b])
        grasp.append((int(xs[len(xs) // 2]), int(ys[len(ys) // 2])))
    patches, meta, sizes = eng.collect_samples(dt, mask=mt, seed=seed, first_frame_index=first, grasp_xy=grasp,
                                               total_score=[0.5 + 0.01 * b for b in range(n)])
    patches = patches.cpu().numpy()
    sets = [(O.collector_tip_points(masks[b]), O.collector_stem_points(masks[b]), O.collector_edge_points(masks[b]))
            for b in range(n)]
    assert sizes.tolist() == [[len(s) for s in sets[b]] for b in range(n)]
    for kind in range(3):
        nq = max(1, max(len(sets[b][kind]) for b in range(n)))
        got = eng.collector_points(kind, np.tile(np.arange(nq), (n, 1)), mask=mt)
        for b in range(n):
            want = sets[b][kind]
            if not want:
                assert (got[b] == -1).all()
            else:
                assert got[b, :len(want)].tolist() == [list(p) for p in want], f"frame {b} kind {kind}"
    trace = []
    for b in range(n):
        s = O.score_maps(masks[b], depth[b], f, cx, cy, "strict")
        want = O.collect_sample(masks[b], depth[b], s, grasp[b], 0.5 + 0.01 * b, O.CollectorRng(seed, first + b),
                                trace=trace)
        valid = [k for k in range(N.SAMPLES_PER_FRAME) if meta[b, k]["valid"]]
        if want is None:
            assert valid == [], f"frame {b}"
            continue
        assert len(valid) == len(want), f"frame {b}"
        for k, w in zip(valid, want):
            m = meta[b, k]
            assert (int(m["kind"]), int(m["label"]), int(m["is_augmented"])) == (w["kind"], w["label"], int(w["is_augmented"]))
            assert (int(m["x"]), int(m["y"])) == tuple(w["grasp_point"]), f"frame {b} slot {k}"
            assert float(m["total_score"]) == w["total_score"]
            _check_patch(patches[b, k], w["patch"], bool(w["is_augmented"]), f"frame {b} slot {k}")
    # the masks are built so that some picks lie too close to the image border: the retry rounds are exercised
    assert any(not ok for (_, _, _, ok) in trace) and any(a > 0 for (a, _, _, _) in trace)


def test_collector_batch_after_process_batch():
    """Frames of one batch get consecutive generator indices; frames without a leaf or without a candidate give nothing."""
    spec = synth.SMALL
    n = 6
    lab, dep = synth.make_batch(spec, CU.CC.CONFIG_SEED, 0, n)
    lab[4] = 0                                                 # no leaf at all
    P = synth.projection_matrix(spec)
    f, cx, cy = P[0, 0], P[0, 2], P[1, 2]
    eng = _engine(n, spec.height, spec.width)
    lt, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda()
    res = eng.process_batch(lt, dt, _cam(P))
    patches, meta, sizes = eng.collect_samples(dt, labels=lt, seed=5, first_frame_index=1000)
    patches = patches.cpu().numpy()
    for b in range(n):
        if res["leaf_id"][b] < 0:
            assert not meta[b]["valid"].any() and sizes[b].tolist() == [0, 0, 0]
            continue
        mask = (lab[b] == res["leaf_id"][b]).astype(np.uint8)
        s = O.score_maps(mask, dep[b], f, cx, cy, "strict")
        if res["n_candidates"][b] == 0:
            assert not meta[b]["valid"].any()
            continue
        g = (int(res["grasp_x"][b]), int(res["grasp_y"][b]))
        total = float(np.max(s["traditional_score"]))
        want = O.collect_sample(mask, dep[b], s, g, total, O.CollectorRng(5, 1000 + b))
        valid = [k for k in range(N.SAMPLES_PER_FRAME) if meta[b, k]["valid"]]
        assert want is not None and len(valid) == len(want)
        for k, w in zip(valid, want):
            m = meta[b, k]
            assert (int(m["x"]), int(m["y"])) == tuple(w["grasp_point"])
            np.testing.assert_allclose(float(m["total_score"]), w["total_score"], rtol=1e-6)
            _check_patch(patches[b, k], w["patch"], bool(w["is_augmented"]), f"frame {b} slot {k}")
    # same call again: same samples (pure function of seed and frame index); another seed: other noise
    p2, m2, _ = eng.collect_samples(dt, labels=lt, seed=5, first_frame_index=1000)
    assert np.array_equal(p2.cpu().numpy(), patches) and np.array_equal(m2, meta)
    p3, _, _ = eng.collect_samples(dt, labels=lt, seed=6, first_frame_index=1000)
    assert not np.array_equal(p3.cpu().numpy()[:, 1, 0], patches[:, 1, 0])
    assert np.array_equal(p3.cpu().numpy()[:, 0], patches[:, 0])


def test_collector_host_class_end_to_end(tmp_path):
    """The reference-shaped collect_sample and the batched entry point write the reference's training_data.pt layout."""
    from leafgrasp_b200 import EnhancedGraspDataCollector
    mask, dep, P, g, _, _ = CU.case_inputs("crafted1")
    H, W = mask.shape
    eng = _engine(2, H, W)
    col = EnhancedGraspDataCollector(resume=False, data_dir=str(tmp_path / "d"), engine=eng, seed=int(g["rng_seed"]))
    col.set_camera_params(P)
    col.frames_seen = int(g["frame_index"])
    ok = col.collect_sample(torch.from_numpy(mask.astype(bool)).cuda(), torch.from_numpy(dep).cuda(), None, {},
                            tuple(g["grasp"].tolist()), float(g["total_score"]))
    assert ok and len(col.samples) == len(g["labels"])
    assert col.stats == {"positive_samples": 1, "augmented_samples": 3, "negative_samples": 3}
    for i, s in enumerate(col.samples):
        assert list(s["grasp_point"]) == g["points"][i].tolist() and s["label"] == int(g["labels"][i])
        got = torch.cat([s["depth_patch"][None], s["mask_patch"][None], s["score_patches"]]).numpy()
        _check_patch(got, g["patches"][i], s["is_augmented"], f"sample {i}")
    # batched: two frames in one call
    m2 = torch.from_numpy(np.stack([mask, mask])).cuda()
    d2 = torch.from_numpy(np.stack([dep, dep])).cuda()
    eng.select_grasp_point(m2, d2, _cam(P))
    assert col.collect_from_engine(d2, mask=m2) == 2
    assert col.stats["positive_samples"] == 3
    col.save_samples()
    data = torch.load(str(tmp_path / "d" / "training_data.pt"))
    assert data["depth_patches"].shape[0] == len(col.samples) and data["score_patches"].shape[1:] == (7, 32, 32)
    # a grasp point too close to the border gives nothing (data_collector.py:83-89)
    assert not col.collect_sample(torch.from_numpy(mask.astype(bool)).cuda(), torch.from_numpy(dep).cuda(), None, {},
                                  (5, 100), 0.5)


def test_collector_full_size_frames():
    """The metric's frame size (1440x1080, 30 leaves): samples of two frames of a batch against the oracle."""
    spec, n = synth.CFG2, 2
    lab, dep = synth.make_batch(spec, CU.CC.CONFIG_SEED, 0, n)
    P = synth.projection_matrix(spec)
    f, cx, cy = P[0, 0], P[0, 2], P[1, 2]
    eng = _engine(n, spec.height, spec.width)
    lt, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda()
    res = eng.process_batch(lt, dt, _cam(P))
    patches, meta, sizes = eng.collect_samples(dt, labels=lt, seed=9, first_frame_index=7)
    patches = patches.cpu().numpy()
    for b in range(n):
        assert res["leaf_id"][b] >= 0 and res["n_candidates"][b] > 0
        mask = (lab[b] == res["leaf_id"][b]).astype(np.uint8)
        s = O.score_maps(mask, dep[b], f, cx, cy, "strict")
        g = (int(res["grasp_x"][b]), int(res["grasp_y"][b]))
        want = O.collect_sample(mask, dep[b], s, g, float(np.max(s["traditional_score"])), O.CollectorRng(9, 7 + b))
        assert sizes[b].tolist() == [len(O.collector_tip_points(mask)), len(O.collector_stem_points(mask)),
                                     len(O.collector_edge_points(mask))]
        valid = [k for k in range(N.SAMPLES_PER_FRAME) if meta[b, k]["valid"]]
        assert want is not None and len(valid) == len(want)
        for k, w in zip(valid, want):
            assert (int(meta[b, k]["x"]), int(meta[b, k]["y"])) == tuple(w["grasp_point"])
            _check_patch(patches[b, k], w["patch"], bool(w["is_augmented"]), f"frame {b} slot {k}")
