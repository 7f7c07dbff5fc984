"""CPU tests of the host side: the C-ABI library loads and exports everything the header declares, struct
layouts agree, BatchNorm folding / weight packing is right, the drop-in module keeps the reference's
state_dict keys, and frame sharding + record gathering work across 2 gloo ranks."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import leafgrasp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from leafgrasp_b200 import _native as N
    h = N.lib()
    header = open(os.path.join(ROOT, "include", "leafgrasp.h")).read()
    declared = set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
    for name in declared:
        assert hasattr(h, name)
    assert h.lg_sizeof_frame_result() == N.FRAME_RESULT.itemsize
    assert h.lg_sizeof_leaf_record() == N.LEAF_RECORD.itemsize


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must refuse, not fall back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from leafgrasp_b200 import GraspEngine, GraspPointCNN, _native as N
    with pytest.raises(N.NativeError):
        GraspEngine(1, 64, 64, 2)
    net = GraspPointCNN().eval()
    with pytest.raises(N.NativeError):
        net(torch.zeros(1, 9, 32, 32))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "leaf-grasping-vision-ml_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "leafgrasp_oracle" not in src and "import oracle" not in src, fn


def test_state_dict_keys_match_reference_layout():
    from leafgrasp_b200 import GraspPointCNN, pack_weights, _native as N
    sd = O.seeded_state_dict(1)
    net = GraspPointCNN(in_channels=9)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    assert pack_weights(net.state_dict()).size == N.lib().lg_cnn_weight_floats()
    with pytest.raises(NotImplementedError):
        GraspPointCNN(in_channels=3)


def _variants():
    import json
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "cnn_variants.json")))
    return meta["variants"]


def test_variant_architectures_keep_the_reference_state_dict_layout():
    """Every architecture of the reference's sweep: same keys and shapes as the reference's own class produced
    (tests/golden/cnn_variants.json), and the packed blob has the size the library expects for that architecture."""
    from leafgrasp_b200 import GraspPointCNN, pack_weights, _native as N
    from leafgrasp_b200.cnn import architecture_of
    lib = N.lib()
    for v in _variants():
        net = GraspPointCNN(in_channels=9, attention_type=v["attention_type"], encoder_filters=v["encoder_filters"])
        mine = {k: list(t.shape) for k, t in net.state_dict().items()}
        assert mine == v["shapes"], (v["attention_type"], v["encoder_filters"])
        sd = O.seeded_state_dict_from_shapes(v["shapes"], v["seed"])
        net.load_state_dict(sd)
        assert architecture_of(sd) == (v["attention_type"], v["encoder_filters"])
        cfg = N.cnn_config(v["attention_type"], v["encoder_filters"])
        assert pack_weights(sd).size == lib.lg_cnn_model_floats(cfg)
    bad = N.cnn_config("spatial", [64, 100])          # 100 channels: not a multiple of 16
    assert lib.lg_cnn_model_floats(bad) == 0


def _folded_forward(folded, x):
    """The folded weights through plain torch ops (float64): what the CUDA kernels compute."""
    y = x.double()
    for l, (w, b) in enumerate(folded["convs"]):
        y = F.relu(F.conv2d(y, torch.from_numpy(w), torch.from_numpy(b), padding=1))
        if l % 2 == 1:
            y = F.max_pool2d(y, 2)
    att = torch.ones_like(y[:, :1])
    if folded["spatial"] is not None:
        aw, ab = folded["spatial"]
        att = torch.sigmoid(F.conv2d(y, torch.from_numpy(aw), torch.from_numpy(ab)))
    catt = torch.ones_like(y[:, :, :1, :1])
    if folded["channel"] is not None:
        (w1, b1), (w2, b2) = folded["channel"]
        h = F.relu(F.conv2d(y.mean(dim=(2, 3), keepdim=True), torch.from_numpy(w1), torch.from_numpy(b1)))
        catt = torch.sigmoid(F.conv2d(h, torch.from_numpy(w2), torch.from_numpy(b2)))
    y = (y * att * catt).mean(dim=(2, 3))
    for k, (w, b) in enumerate(folded["fcs"]):
        y = F.linear(y, torch.from_numpy(w), torch.from_numpy(b))
        if k != 3:
            y = F.relu(y)
    return y.float()


def test_batchnorm_folding_is_exact_enough():
    """Run the folded weights through plain torch ops and compare with the unfolded forward: the default architecture
    against the oracle, the variants against the logits the reference's own class produced (golden)."""
    from leafgrasp_b200 import fold_batchnorm
    sd = O.seeded_state_dict(5)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 9, 32, 32, generator=g)
    np.testing.assert_allclose(_folded_forward(fold_batchnorm(sd), x).numpy(), O.cnn_forward(sd, x).numpy(), atol=2e-5, rtol=1e-5)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "cnn_variants.npz"))
    xv = torch.from_numpy(gold["x"])
    for i, v in enumerate(_variants()):
        sdv = O.seeded_state_dict_from_shapes(v["shapes"], v["seed"])
        y = _folded_forward(fold_batchnorm(sdv), xv).reshape(-1).numpy()
        np.testing.assert_allclose(y, gold[f"logits_{i}"], atol=5e-5, rtol=1e-4, err_msg=str((v["attention_type"], v["encoder_filters"])))


def test_ellipse_rows_restated_equals_opencv():
    """csrc/lg_api.cu:ellipse_rows restates cv2.getStructuringElement(MORPH_ELLIPSE); same formula here."""
    import cv2
    for n in (30, 31):
        se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (n, n))
        r = c = n // 2
        for i in range(n):
            dy = i - r
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) / (r * r))))
            a, b = max(c - dx, 0), min(c + dx + 1, n)
            row = np.zeros(n, np.uint8)
            row[a:b] = 1
            np.testing.assert_array_equal(row, se[i])


def test_wire_format_round_trip():
    """msg/masks.msg uint16[] / msg/depth.msg float32[] -> tensors as mask_callback / depth_callback build them
    (leaf_grasp_node_v3.py:185-205), and the published result string (:168-173)."""
    from leafgrasp_b200 import wire, _native as N
    H, W = 6, 9
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 40, H * W).astype(np.uint16)
    dep = rng.random(H * W).astype(np.float32)

    class Msg:            # what rospy hands to the callbacks: a tuple of Python numbers
        def __init__(self, data):
            self.imageData = tuple(data.tolist())

    ref_mask = torch.tensor(np.array(Msg(ids).imageData, dtype=np.int16)).reshape(H, W)          # the reference's lines
    ref_depth = torch.tensor(np.array(Msg(dep).imageData, dtype=np.float32)).reshape(H, W)
    assert np.array_equal(wire.mask_from_wire(Msg(ids), H, W), ref_mask.numpy())
    assert np.array_equal(wire.depth_from_wire(Msg(dep), H, W), ref_depth.numpy())
    lab, d = wire.stage_frames([Msg(ids), ids], [Msg(dep), dep], H, W, pin=False)
    assert lab.dtype == torch.int16 and d.dtype == torch.float32 and lab.shape == (2, H, W)
    assert torch.equal(lab[0], ref_mask) and torch.equal(lab[1], ref_mask) and torch.equal(d[1], ref_depth)
    with pytest.raises(ValueError):
        wire.mask_from_wire(ids[:-1], H, W)
    g2, g3, pre = (734, 608), (0.01, -0.02, 0.45), (0.011, -0.021, 0.45)
    assert wire.format_result(g2, g3, pre) == f"{g2[0]},{g2[1]},{g3[0]},{g3[1]},{g3[2]},{pre[0]},{pre[1]},{pre[2]}"
    assert wire.format_result(g2, g3, None) == "734,608,0.01,-0.02,0.45"
    rec = np.zeros(1, dtype=N.FRAME_RESULT)[0]
    rec["leaf_id"] = -1
    assert wire.format_frame_result(rec) is None
    rec["leaf_id"], rec["n_candidates"], rec["grasp_x"], rec["grasp_y"] = 3, 20, 734, 608
    rec["grasp_3d"] = g3
    rec["pre_grasp"] = pre
    assert wire.format_frame_result(rec) == wire.format_result(g2, g3, pre)


def test_records_from_result_buffer_matches_numpy_view():
    """The on-device record builder reads the same bytes as the NumPy structured view of lg_frame_result."""
    from leafgrasp_b200 import _native as N, dist as lgd
    rng = np.random.default_rng(3)
    rec = np.zeros(5, dtype=N.FRAME_RESULT)
    rec["cand_x"] = rng.integers(0, 1440, (5, 20)); rec["cand_y"] = rng.integers(0, 1080, (5, 20))
    rec["trad"] = rng.random((5, 20)); rec["ml"] = rng.random((5, 20)); rec["ml"][2, 7:] = np.nan
    rec["n_candidates"] = [20, 0, 20, 13, 1]
    buf = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
    a = lgd.records_from_result_buffer(buf, 5)
    b = lgd.records_from_results(rec, "cpu")
    assert a.shape == (5, 20, 4) and torch.equal(a, b)
    assert torch.equal(a[3, 13:], torch.tensor([-1.0, -1.0, 0.0, 0.0]).expand(7, 4))     # unused slots
    assert a[3, 12, 0] == float(rec["cand_x"][3, 12]) and a[2, 9, 3] == 0.0               # NaN ML score -> 0


def _dist_worker(rank, world, port, out):
    import torch.distributed as dist
    from leafgrasp_b200 import dist as lgd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = lgd.shard_range(10, rank, world)
    rec = torch.zeros(hi - lo, 20, 4)
    rec[:, :, 0] = torch.arange(lo, hi).float()[:, None]
    allrec = lgd.gather_candidate_records(rec, 10)
    # the benchmark's aggregation: equal blocks, one asynchronous all-gather after the last step
    mine = torch.full((3, 20, 4), float(rank))
    agg, work = lgd.gather_records_async(mine)
    if work is not None:
        work.wait()
    if rank == 0:
        out.put((allrec[:, 0, 0].tolist(), agg[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_frame_sharding_and_gather_two_ranks():
    import torch.multiprocessing as mp
    from leafgrasp_b200 import dist as lgd
    assert lgd.shard_range(10, 0, 2) == (0, 5) and lgd.shard_range(10, 1, 2) == (5, 10)
    assert lgd.shard_range(7, 0, 3) == (0, 3) and lgd.shard_range(7, 2, 3) == (5, 7)
    covered = []
    for r in range(8):
        lo, hi = lgd.shard_range(8192, r, 8)
        covered += list(range(lo, hi))
    assert covered == list(range(8192))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == [float(i) for i in range(10)]
    assert got[1] == [0.0] * 3 + [1.0] * 3


def test_benchmark_weights_equal_the_oracle_weights():
    """bench.py / smoke() take their random-init model from the product package (synth.seeded_state_dict); the golden
    vectors were made with the oracle's: the two generators must produce the same tensors."""
    import leafgrasp_oracle as O
    from leafgrasp_b200 import synth
    a, b = synth.seeded_state_dict(1234), O.seeded_state_dict(1234)
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_process_batch_host_validates_its_buffers():
    """The throughput entry point hands raw host pointers to the library: dtype, shape, device and contiguity are
    checked first (a wrong buffer would otherwise be read out of bounds)."""
    from leafgrasp_b200.pipeline import GraspEngine
    eng = GraspEngine.__new__(GraspEngine)          # no device needed for the checks
    eng.B, eng.H, eng.W = 4, 6, 8
    ok = torch.zeros(2, 6, 8, dtype=torch.int16)
    assert eng._host_frames(ok, torch.int16, "labels") is ok
    assert eng._host_frames(np.zeros((6, 8), np.float32), torch.float32, "depth").shape == (1, 6, 8)
    for bad in (torch.zeros(2, 6, 8, dtype=torch.int32), torch.zeros(2, 6, 9, dtype=torch.int16),
                torch.zeros(5, 6, 8, dtype=torch.int16), torch.zeros(2, 6, 16, dtype=torch.int16)[:, :, ::2]):
        with pytest.raises(ValueError):
            eng._host_frames(bad, torch.int16, "labels")


# ---------------------------------------------------------------------------------------------------
# training-sample collector: host side (data_collector.py), no device needed
# ---------------------------------------------------------------------------------------------------
def _fake_collect(n, rng):
    from leafgrasp_b200 import _native as N
    patches = rng.random((n, N.SAMPLES_PER_FRAME, 9, 32, 32)).astype(np.float32)
    meta = np.zeros((n, N.SAMPLES_PER_FRAME), dtype=N.SAMPLE_META)
    for b in range(n):
        pos_ok = b != 1                                     # frame 1: collect_sample returned False
        for k in range(N.SAMPLES_PER_FRAME):
            m = meta[b, k]
            m["kind"] = k
            if not pos_ok or (k == 6 and b % 2 == 0):       # some frames find only two negatives
                continue
            m["valid"], m["label"], m["is_augmented"] = 1, int(k < 4), int(1 <= k <= 3)
            m["x"], m["y"], m["total_score"] = 100 + b, 50 + k, 0.5 if k < 4 else 0.0
    return patches, meta


def test_collector_host_bookkeeping_and_file_format(tmp_path):
    from leafgrasp_b200 import EnhancedGraspDataCollector, _native as N
    assert N.SAMPLE_META.itemsize == 32
    rng = np.random.default_rng(0)
    col = EnhancedGraspDataCollector(resume=False, data_dir=str(tmp_path / "d"))
    patches, meta = _fake_collect(5, rng)
    assert col.collect_batch(patches, meta) == 4
    assert col.frames_seen == 5
    assert col.stats == {"positive_samples": 4, "augmented_samples": 12, "negative_samples": 4 * 3 - 3}
    assert len(col.samples) == 4 * 7 - 3
    s0 = col.samples[0]
    assert set(s0) == {"depth_patch", "mask_patch", "score_patches", "total_score", "grasp_point", "label", "is_augmented"}
    assert s0["score_patches"].shape == (7, 32, 32) and s0["grasp_point"] == (100, 50) and s0["label"] == 1
    np.testing.assert_array_equal(s0["depth_patch"].numpy(), patches[0, 0, 0])
    np.testing.assert_array_equal(col.samples[1]["score_patches"].numpy(), patches[0, 1, 2:])
    col.save_samples()
    data = torch.load(str(tmp_path / "d" / "training_data.pt"))
    # the layout ml_grasp_optimizer/dataset.py reads (data_collector.py:515-523)
    assert set(data) == {"depth_patches", "mask_patches", "score_patches", "labels", "total_scores", "grasp_points",
                         "is_augmented"}
    assert data["depth_patches"].shape == (25, 32, 32) and data["score_patches"].shape == (25, 7, 32, 32)
    assert data["grasp_points"].shape == (25, 2) and data["is_augmented"].dtype == torch.bool
    text = open(tmp_path / "d" / "collection_metadata.txt").read()
    assert "Original positive samples: 4" in text and "Total samples: 25" in text
    assert open(tmp_path / "d" / "collection_progress.txt").read() == "last_frame: 4\n"
    # resume picks the samples and the counters up again (data_collector.py:42-81)
    col2 = EnhancedGraspDataCollector(resume=True, data_dir=str(tmp_path / "d"))
    assert col2.stats == col.stats and len(col2.samples) == 25 and col2.frames_seen == 4
    assert col2.samples[3]["grasp_point"] == col.samples[3]["grasp_point"]
    # resume=False clears the directory (:19-21)
    col3 = EnhancedGraspDataCollector(resume=False, data_dir=str(tmp_path / "d"))
    assert col3.samples == [] and not os.path.exists(tmp_path / "d" / "training_data.pt")
    with pytest.raises(Exception):
        col3.collect_sample(torch.zeros(4, 4, dtype=torch.bool), torch.zeros(4, 4), None, {}, (1, 1), 0.5)   # no engine


def _rle_reference(lab):
    """NumPy statement of the run format (include/leafgrasp.h: lg_rle_encode_labels)."""
    H, W = lab.shape
    u = lab.view(np.uint16).astype(np.uint32)
    start = np.ones((H, W), dtype=bool)
    start[:, 1:] = u[:, 1:] != u[:, :-1]
    ys, xs = np.nonzero(start)
    runs = xs.astype(np.uint32) | (u[ys, xs] << 16)
    rowoff = np.concatenate([[0], np.cumsum(start.sum(axis=1))]).astype(np.uint32)
    return runs, rowoff


def _rle_expand(runs, rowoff, H, W):
    out = np.empty((H, W), dtype=np.uint16)
    for y in range(H):
        r = runs[rowoff[y]:rowoff[y + 1]]
        x0 = (r & 0xFFFF).astype(np.int64)
        x1 = np.append(x0[1:], W)
        out[y] = np.repeat((r >> 16).astype(np.uint16), x1 - x0)
    return out.view(np.int16)


@pytest.mark.parametrize("shape", [(1, 1), (3, 15), (5, 16), (7, 17), (9, 33), (16, 47), (8, 64), (37, 1440), (360, 480), (20, 4099)])
def test_host_label_encoder_versions_agree_and_invert(shape):
    """The label encoder of lg_process_batch_host (csrc/lg_rle_host.cpp) is host code, so it is checked here: every
    instruction-set version the CPU offers writes exactly the words of the NumPy statement of the format, expanding the runs
    gives the image back (negative ids and noise included), and a run table that is too small is reported, not overrun."""
    import ctypes as C
    from leafgrasp_b200 import _native as N, synth
    h = N.lib()
    H, W = shape
    rng = np.random.default_rng(H * 10007 + W)
    if (H, W) == (360, 480):
        images = [synth.make_frame(synth.SMALL, 7, k)[0] for k in range(3)]
    else:
        blocky = np.repeat(np.repeat(rng.integers(0, 40, size=(-(-H // 4), -(-W // 37))), 4, axis=0), 37, axis=1)[:H, :W]
        images = [blocky.astype(np.int16),
                  rng.integers(-3, 3, size=(H, W)).astype(np.int16),                    # noise, negative ids
                  np.full((H, W), 32767, dtype=np.int16),
                  (np.arange(W)[None, :] % 2 * -32768 + np.zeros((H, 1))).astype(np.int16)]   # a run per pixel
    best = h.lg_rle_host_isa()
    assert 0 <= best <= 2
    for lab in images:
        lab = np.ascontiguousarray(lab)
        want_runs, want_off = _rle_reference(lab)
        n = len(want_runs)
        assert np.array_equal(_rle_expand(want_runs, want_off, H, W), lab)
        for isa in [-1] + list(range(best + 1)):
            runs = np.full(n + 8, 0xDEADBEEF, dtype=np.uint32)
            off = np.full(H + 2, 0xDEADBEEF, dtype=np.uint32)
            got = h.lg_rle_encode_labels(lab.ctypes.data_as(C.c_void_p), H, W, runs.ctypes.data_as(C.c_void_p), n,
                                         off.ctypes.data_as(C.c_void_p), isa)
            assert got == n, (isa, got, n)
            assert np.array_equal(runs[:n], want_runs) and np.all(runs[n:] == 0xDEADBEEF), isa
            assert np.array_equal(off[:H + 1], want_off) and off[H + 1] == 0xDEADBEEF, isa
            if n > 1:                                   # one word too few: overflow, and nothing written past the table
                runs[:] = 0xDEADBEEF
                got = h.lg_rle_encode_labels(lab.ctypes.data_as(C.c_void_p), H, W, runs.ctypes.data_as(C.c_void_p), n - 1,
                                             off.ctypes.data_as(C.c_void_p), isa)
                assert got == 0xFFFFFFFF and np.all(runs[n - 1:] == 0xDEADBEEF), isa
    bad = np.zeros((2, 2), dtype=np.int16)
    assert h.lg_rle_encode_labels(None, 2, 2, None, 4, None, -1) == 0xFFFFFFFF
    assert h.lg_rle_encode_labels(bad.ctypes.data_as(C.c_void_p), 2, 70000, bad.ctypes.data_as(C.c_void_p), 4,
                                  bad.ctypes.data_as(C.c_void_p), -1) == 0xFFFFFFFF
