#!/usr/bin/env python
"""Benchmark of the grasp-selection hot path (BASELINE.json metric: frames/s/GPU at 1440x1080, 30 leaves).

    python bench.py [--config cfg2] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cnn bf16|fp32]

--config picks one of BASELINE.json's five configurations (default cfg2, the one the metric is quoted on):
  cfg1  one 1440x1080 frame with 10 leaves through the drop-in classes (select_optimal_leaf + select_grasp_point): latency
  cfg2  256 frames x 30 leaves per GPU and step, the whole path                       (the headline line)
  cfg3  3840x2160 frames with 100 overlapping leaves
  cfg4  GraspPointCNN only: 65 536 patches, tcgen05 bf16 against the fp32 CUDA-core kernels
  cfg5  8192 frames per step sharded over the N GPUs (strong scaling), ONE NCCL all-gather of the candidate records
One "step" = one pass of the path over one batch of synthetic input.  Prints ONE JSON line (rank 0):
  value     whole-job throughput with the batch resident in HBM, timed on the device with CUDA events,
            barrier + synchronize on both sides, max over ranks;
  e2e       the same through the host-buffer entry point: pinned host buffers in, host records out, copies timed;
  roofline  the dominant kernel of the step (per-stage CUDA events recorded inside the library over the timed region):
            algorithmic bytes per launch / its mean device time, against MEASURED_PEAKS.json;
  cpu_baseline  the reference's CPU path timed on this box's host cores (bounded sample);
  consistency   every GPU arm carries frames of the golden set (tests/golden/, produced by the unmodified reference) inside
            the timed batch and compares leaf id, positive-key candidates and grasp pixel with the reference's answers.
The default run (cfg2, one GPU) appends short cfg1 / cfg3 / cfg4 measurements under "extra".
--impl reference times the reference's CPU path alone: the unmodified reference modules when build() staged them under
baseline/_ref (kind "reference"), else the oracle port (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "grasp_selection_frames_per_sec"
UNIT = "frames/s"
WORKLOADS = {
    "cfg1": "cfg1: one 1440x1080 frame, 10 leaves, drop-in classes (select_optimal_leaf + select_grasp_point), latency",
    "cfg2": "cfg2: 1440x1080 frames, 30 leaves, full path (leaf selection + score maps + top-20 + patches + GraspPointCNN + fusion)",
    "cfg3": "cfg3: 3840x2160 frames, 100 overlapping leaves, full path",
    "cfg4": "cfg4: GraspPointCNN only, 65536 patches [9,32,32]",
    "cfg5": "cfg5: 8192 1440x1080 frames x 30 leaves per step, frame-sharded over the GPUs, one all-gather of candidate records",
}
CONFIG_SEED = 11       # the benchmark's own frames
GOLDEN_SEED = 7        # frames the reference's answers are stored for (tests/golden)
CNN_SEED = 1234
STAGES = ["", "leaf_stats", "scatter", "median", "edt_columns", "edt_rows", "select", "chamfer", "orientation",
          "score_maps", "candidates", "patches", "cnn", "fuse"]
GOLD = os.path.join(ROOT, "tests", "golden")
CNN_FLOP = 312.83e6    # per patch


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "tf_burst": float(d["bf16_tflops"]),
                "tf_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1600.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------------------
def _frame_job(job):
    from leafgrasp_b200 import synth
    spec_name, seed, idx = job
    return synth.make_frame(getattr(synth, spec_name), seed, idx)


def make_frames(spec_name, jobs, workers):
    """[(seed, index)] -> labels int16 [n,H,W], depth float32 [n,H,W]; generated by a pool of forked workers
    (host-side input generation, before CUDA is touched; never timed)."""
    from leafgrasp_b200 import synth
    spec = getattr(synth, spec_name)
    lab = np.empty((len(jobs), spec.height, spec.width), np.int16)
    dep = np.empty((len(jobs), spec.height, spec.width), np.float32)
    work = [(spec_name, s, i) for s, i in jobs]
    if workers > 1 and len(jobs) > 2:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            for k, (l, d) in enumerate(pool.imap(_frame_job, work, chunksize=1)):
                lab[k], dep[k] = l, d
    else:
        for k, j in enumerate(work):
            lab[k], dep[k] = _frame_job(j)
    return lab, dep


def golden_frames(spec_name):
    """[(index, npz)] of the golden frames stored for this spec."""
    out = []
    for meta_name in ("golden_meta.json", "golden_meta_r2.json"):
        p = os.path.join(GOLD, meta_name)
        if not os.path.exists(p):
            continue
        for fr in json.load(open(p))["frames"]:
            if fr["spec"] == spec_name:
                out.append((int(fr["index"]), np.load(os.path.join(GOLD, fr["file"]))))
    return sorted(out, key=lambda t: t[0])


def records_equal(a, b):
    """Field-wise equality of two lg_frame_result arrays (NaN == NaN; struct padding is not compared)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return False
    for f in a.dtype.names:
        x, y = a[f], b[f]
        if not np.array_equal(x, y, equal_nan=x.dtype.kind == "f"):
            return False
    return True


def check_golden(records, gold, where, compare_grasp):
    """records[where[k]] against the reference's stored answers for golden frame k."""
    res = {"frames": len(gold), "leaf_id_identical": 0, "positive_candidates_identical": 0, "grasp_identical": 0,
           "grasp_compared": bool(compare_grasp)}
    for (idx, g), pos in zip(gold, where):
        r = records[pos]
        res["leaf_id_identical"] += int(int(r["leaf_id"]) == int(g["leaf_id"]))
        if int(g["leaf_id"]) < 0:
            res["positive_candidates_identical"] += 1
            res["grasp_identical"] += 1
            continue
        npos = int(g["n_positive"])
        cand = np.stack([r["cand_x"], r["cand_y"]], axis=1)
        res["positive_candidates_identical"] += int(int(r["n_positive"]) == npos and
                                                    np.array_equal(cand[:npos], g["candidates"][:npos]))
        res["grasp_identical"] += int((int(r["grasp_x"]), int(r["grasp_y"])) == tuple(int(v) for v in g["grasp_2d"]))
    res["identical"] = bool(res["leaf_id_identical"] == len(gold) and res["positive_candidates_identical"] == len(gold)
                            and (not compare_grasp or res["grasp_identical"] == len(gold)))
    return res


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's path on the host cores
# ------------------------------------------------------------------------------------------------------
_W = {}
REF_STAGE = os.path.join(ROOT, "baseline", "_ref")


def reference_staged():
    return os.path.exists(os.path.join(REF_STAGE, "scripts", "utils", "grasp_point_selector.py"))


def _cpu_init(spec_name, seed, use_real):
    import cv2
    import torch
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(GOLD)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import leafgrasp_oracle as O
    from leafgrasp_b200 import synth
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    _W["O"], _W["synth"] = O, synth
    _W["spec"] = getattr(synth, spec_name)
    _W["P"] = synth.projection_matrix(_W["spec"])
    _W["sd"] = O.seeded_state_dict(CNN_SEED)
    _W["real"] = None
    if use_real:
        import warnings
        warnings.filterwarnings("ignore")
        import ref_harness
        ref_harness.REFERENCE_ROOT = REF_STAGE
        OLS, GPS, IP, CNN = ref_harness.load()
        dev = torch.device("cpu")
        ols, gps = OLS(dev), GPS(dev)
        ols.set_camera_params(_W["P"])
        gps.set_camera_params(_W["P"])
        net = CNN(in_channels=9)
        net.load_state_dict(_W["sd"])
        net.eval()
        gps.ml_predictor = net
        _W["real"] = (ols, gps, IP(_W["spec"].height, _W["spec"].width, 21, 5))
    # one seeded frame per worker process, generated before any timing
    _W["frame"] = synth.make_frame(_W["spec"], seed, os.getpid() % 4096)


def _cpu_frame(idx):
    import torch
    lab, dep = _W["frame"]
    t = time.perf_counter()
    if _W["real"] is not None:      # leaf_grasp_node_v3.py:110-119 on the unmodified modules
        ols, gps, ip = _W["real"]
        mt, dt = torch.from_numpy(lab), torch.from_numpy(dep)
        leaf = ols.select_optimal_leaf(mt, dt)
        if leaf is not None:
            gps.select_grasp_point(mt == leaf, dt, ip)
    else:
        _W["O"].process_frame(lab, dep, _W["P"], _W["sd"], arith="reference")
    return time.perf_counter() - t


def cpu_arm(spec_name, steps, warmup, workers=None):
    """Each step: `workers` frames, one per worker process (1 thread each) -> frames/s over all host cores."""
    import multiprocessing as mp
    workers = workers or max(1, min(os.cpu_count() or 1, 32))
    use_real = reference_staged()
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(spec_name, CONFIG_SEED, use_real)) as pool:
        frames = list(range(workers))
        for _ in range(max(1, warmup)):
            pool.map(_cpu_frame, frames, chunksize=1)
        t0 = time.perf_counter()
        per = []
        for _ in range(steps):
            per += pool.map(_cpu_frame, frames, chunksize=1)
        wall = time.perf_counter() - t0
    what = "the unmodified reference modules (baseline/_ref)" if use_real else "the oracle port"
    return {"value": workers * steps / wall, "unit": UNIT, "cores": workers, "kind": "reference" if use_real else "port",
            "sample": f"{steps} steps x {workers} {spec_name.lower()} frames, one process per core (1 thread each) running {what}; "
                      f"mean {statistics.mean(per):.2f} s/frame/core",
            "ms_per_step": wall / steps * 1e3, "frames_per_step": workers}


def cpu_cnn_arm(steps, warmup, n_patches=2048):
    """cfg4 on the host: the reference's GraspPointCNN (torch CPU, all threads) on a bounded sample of patches."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import leafgrasp_oracle as O
    sd = O.seeded_state_dict(CNN_SEED)
    g = torch.Generator().manual_seed(4)
    x = torch.rand(n_patches, 9, 32, 32, generator=g)
    x[:, 1] = (x[:, 1] > 0.5).float()
    fwd = lambda: O.cnn_forward(sd, x)
    kind = "port"
    if reference_staged():
        sys.path.insert(0, GOLD)
        import ref_harness
        ref_harness.REFERENCE_ROOT = REF_STAGE
        net = ref_harness.load()[3](in_channels=9)
        net.load_state_dict(sd)
        net.eval()
        fwd = lambda: net(x)
        kind = "reference"
    with torch.no_grad():
        for _ in range(max(1, warmup)):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        wall = time.perf_counter() - t0
    return {"value": n_patches * steps / wall, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} steps x {n_patches} patches, torch CPU fp32, {torch.get_num_threads()} threads",
            "ms_per_step": wall / steps * 1e3}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.config == "cfg4":
        r = cpu_cnn_arm(args.steps, args.warmup)
        metric, unit = "cnn_patches_per_sec", "patches/s"
    else:
        spec_name = {"cfg1": "CFG1", "cfg3": "CFG3"}.get(args.config, "CFG2")
        r = cpu_arm(spec_name, args.steps, args.warmup, workers=1 if args.config == "cfg1" else None)
        metric, unit = METRIC, UNIT
    line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.config == "cfg5" else "weak", "vs_baseline": None,
            "dtype": "f64/f32 (NumPy, OpenCV, torch CPU)", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "sample": r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm: frame workloads (cfg2, cfg3, cfg5)
# ------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch, self.dist = torch, dist
        self.dev = None

    def init_cuda(self):
        torch = self.torch
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            self.dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_frames(args, D, config, spec_name, B, chunks_per_step, unique, want_cpu, gold_limit):
    """B frames per lg_process_batch call, chunks_per_step calls per step and rank."""
    import ctypes as C
    from leafgrasp_b200 import hostmem
    torch = D.torch
    world, rank = D.world, D.rank
    cpus = None if args.no_bind else hostmem.bind_to_gpu(D.local)      # before the pinned buffers are allocated
    from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth
    from leafgrasp_b200 import _native as N
    from leafgrasp_b200 import dist as lgd
    spec = getattr(synth, spec_name)
    H, W = spec.height, spec.width
    P = H * W
    cam = camera_from_projection(synth.projection_matrix(spec))
    use_bf16 = args.cnn == "bf16"

    # ---- synthetic batch: `unique` distinct seeded frames (tiled to B when B is larger); on rank 0 the first frames are
    # the golden ones (seed 7), whose answers from the unmodified reference are stored in tests/golden
    gold = golden_frames(spec_name)[:gold_limit] if rank == 0 else []
    unique = max(1, min(B, unique))
    jobs = [(GOLDEN_SEED, idx) for idx, _ in gold][:unique]
    gold = gold[:len(jobs)]
    jobs += [(CONFIG_SEED, rank * unique + k) for k in range(unique - len(jobs))]
    workers = max(1, min(16, len(os.sched_getaffinity(0))))
    lab_u, dep_u = make_frames(spec_name, jobs, workers)
    reps = (B + unique - 1) // unique
    lab_h = hostmem.pinned_like(np.tile(lab_u, (reps, 1, 1))[:B] if reps > 1 else lab_u)
    dep_h = hostmem.pinned_like(np.tile(dep_u, (reps, 1, 1))[:B] if reps > 1 else dep_u)
    D.init_cuda()
    dev = D.dev
    lab_d, dep_d = lab_h.to(dev), dep_h.to(dev)
    pinned_ok = (N.lib().lg_host_memory_is_pinned(C.c_void_p(lab_h.data_ptr())) == 1 and
                 N.lib().lg_host_memory_is_pinned(C.c_void_p(dep_h.data_ptr())) == 1)

    eng = GraspEngine(B, H, W, 128, device=dev, lanes=args.lanes)
    n_prof = eng.lane_split(B)[0][1]          # frames the profiled (main) context handles per call
    eng.set_cnn_weights(pack_weights(synth.seeded_state_dict(CNN_SEED)))
    eng.set_patch_export(False)      # throughput mode: the gather writes the CNN's input layout, no float32 patch tensor
    lib = N.lib()
    steps, cps = args.steps, chunks_per_step
    # candidate records of every frame of the timed region, written by the fusion kernel; ONE all-gather at the end
    rec_all = torch.zeros(steps * cps * B, N.TOP_K, 4, dtype=torch.float32, device=dev)
    rec_gathered = torch.empty(world * steps * cps * B, N.TOP_K, 4, dtype=torch.float32, device=dev) if world > 1 else None

    def run_steps(n_steps, host):
        res = None
        for s in range(n_steps):
            for k in range(cps):
                at = ((s * cps + k) % (steps * cps)) * B
                eng.set_record_output(rec_all[at:at + B])
                if host:
                    res = eng.process_batch_host(lab_h, dep_h, cam, use_bf16)
                else:
                    res = eng.process_batch(lab_d, dep_d, cam, use_bf16, sync=False)
        if world > 1:
            _, work = lgd.gather_records_async(rec_all, rec_gathered)
            work.wait()
        return res

    run_steps(args.warmup, False)
    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------
    lib.lg_set_profiling(eng._ctx, 1)
    sampler = ClockSampler(D.local)
    launches0 = lib.lg_launch_count()
    D.barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run_steps(steps, False)
    e1.record()
    D.barrier()
    clocks = sampler.stop()
    launches = int(lib.lg_launch_count() - launches0)
    ms = D.max_over_ranks(e0.elapsed_time(e1))
    buf = (C.c_float * 14)()
    calls = C.c_int(0)
    lib.lg_stage_times_mean(eng._ctx, buf, 14, C.byref(calls))
    stage_ms = np.array(list(buf))
    lib.lg_set_profiling(eng._ctx, 0)
    records = np.frombuffer(res.cpu().numpy().tobytes(), dtype=N.FRAME_RESULT)
    rec_dev = rec_all[-B:].cpu().numpy()

    # ---- timed region 2: end to end through the host-buffer entry point ----------------------------
    run_steps(2, True)
    D.barrier()
    t0 = time.perf_counter()
    out_host = run_steps(steps, True)
    D.barrier()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0)
    h2d_call, d2h_call = eng.host_call_bytes()        # what one host call really moved (labels cross the link run-length encoded)
    # ---- what the box allows end to end: the same pinned buffers copied to the GPUs by all ranks at once, nothing else
    # running (tools/h2d_ceiling.py measures the same thing standalone, with a solo leg beside it)
    lab_d.copy_(lab_h, non_blocking=True); dep_d.copy_(dep_h, non_blocking=True)
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        lab_d.copy_(lab_h, non_blocking=True); dep_d.copy_(dep_h, non_blocking=True)
    torch.cuda.synchronize()
    own_s = time.perf_counter() - t0
    ceil_s = D.max_over_ranks(own_s)
    h2d_ceiling_gbs = 3 * B * P * 6 * world / ceil_s / 1e9            # equal shards: the slowest link sets the pace
    h2d_links_gbs = D.sum_over_ranks(3 * B * P * 6 / own_s / 1e9)      # every link at its own pace
    if rank != 0:
        eng.close()
        return None

    # ---- roofline of the dominant stage -------------------------------------------------------------
    # stage_ms: CUDA events recorded inside the library during the timed region, mean over its calls (stages on the
    # library's side stream overlap the others).  The dominant kernel is picked from one extra, untimed pass with the
    # overlap switched off, i.e. by the kernels' own durations; `achieved` uses its duration inside the timed region.
    pk = peaks()
    eng.set_record_output(None)
    eng.set_overlap(False)
    eng.lanes_active = False
    lib.lg_set_profiling(eng._ctx, 1)
    eng.process_batch(lab_d, dep_d, cam, use_bf16, sync=False)     # rank 0 only: no collective here
    buf = (C.c_float * 14)()
    lib.lg_stage_times(eng._ctx, buf, 14)
    serial_ms = np.array(list(buf))
    lib.lg_set_profiling(eng._ctx, 0)
    eng.set_overlap(True)
    eng.lanes_active = True
    reg = records["region"].astype(np.int64)
    bbox_px = float(np.mean(np.maximum(reg[:, 2] - reg[:, 0], 0) * np.maximum(reg[:, 3] - reg[:, 1], 0)))
    rect_px = float(np.mean(np.maximum(np.minimum(reg[:, 2] + 16, W) - np.maximum(reg[:, 0] - 16, 0), 0) *
                            np.maximum(np.minimum(reg[:, 3] + 16, H) - np.maximum(reg[:, 1] - 16, 0), 0)))
    leaf_px = float(np.mean((lab_u > 0).sum(axis=(1, 2))))
    n_patches = float(np.mean(records["ml_valid"].sum(axis=1)))
    alg_bytes = {   # per frame, SURVEY.md 8(d): every input read once + every output written once
        "leaf_stats": 6 * P,                       # K0: labels 2 B + depth 4 B per pixel, outputs O(leaves)
        "scatter": 6 * P + 4 * leaf_px, "median": 4 * leaf_px, "edt_columns": 0,
        "edt_rows": 1 * P,                          # K1 on the union mask, arg-max only: one byte-equivalent per pixel in
        "select": 0,
        "chamfer": 5 * bbox_px,                     # K1 inside transform on the leaf rectangle: 1 B in + 4 B out
        "orientation": 1 * bbox_px,
        "score_maps": 19 * rect_px,                 # K2 throughput mode: read 14, write trad f32 + valid u8
        "candidates": 5 * rect_px,                  # K3: trad 4 + valid 1
        "patches": n_patches * 9 * 1024 * 6,        # K4: 4 B read + 2 B written per patch element
        "fuse": 760}
    top = int(np.argmax(serial_ms))
    name = STAGES[top]
    traffic = None      # DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), same batch only
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(name)
        if ent and int(ent.get("frames", 0)) > 0 and ent.get("config", "cfg2") == ("cfg3" if config == "cfg3" else "cfg2"):
            traffic = float(ent["dram_bytes_per_launch"]) * n_prof / int(ent["frames"])
    except Exception:  # noqa: BLE001
        pass
    if name == "cnn":
        ach = CNN_FLOP * n_patches * n_prof / (stage_ms[top] * 1e-3) / 1e12
        roof = {"kernel": "cnn", "bound": "tensor", "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": ach / pk["tf_burst"], "traffic": traffic, "peak_source": pk["source"] + " (burst: the stage lasts ~1 ms)"}
    else:
        ach = alg_bytes[name] * n_prof / (stage_ms[top] * 1e-3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": traffic, "peak_source": pk["source"],
                "alg_bytes_per_launch": alg_bytes[name] * n_prof}
    roof["stage_ms"] = {STAGES[i]: round(float(stage_ms[i]), 4) for i in range(1, len(STAGES))}
    roof["stage_ms_calls_averaged"] = int(calls.value)
    roof["stage_ms_frames"] = n_prof      # stage_ms: the main context's share of the batch (lanes run side by side)
    roof["stage_ms_serial"] = {STAGES[i]: round(float(serial_ms[i]), 4) for i in range(1, len(STAGES))}
    roof["stage_gbs"] = {k: round(alg_bytes[k] * B / (serial_ms[STAGES.index(k)] * 1e-3) / 1e9, 1)
                         for k in alg_bytes if serial_ms[STAGES.index(k)] > 0 and alg_bytes[k] > 0}
    roof["stage_frac_of_hbm_peak"] = {k: round(v / pk["hbm"], 3) for k, v in roof["stage_gbs"].items()}
    i_cnn = STAGES.index("cnn")
    roof["cnn_tflops"] = round(CNN_FLOP * n_patches * B / (serial_ms[i_cnn] * 1e-3) / 1e12, 1) if serial_ms[i_cnn] > 0 else None
    roof["cnn_frac_of_burst_peak"] = round(roof["cnn_tflops"] / pk["tf_burst"], 3) if roof["cnn_tflops"] else None
    step_s = ms / steps * 1e-3
    roof["whole_step_gbs"] = round((45 * P) * B * cps / step_s / 1e9, 1)   # 45 B/px, SURVEY.md 8d

    # ---- parity of the timed batches -------------------------------------------------------------------
    # (a) the golden frames inside the device-resident batch and inside the host-call batch, against the reference's answers
    # (b) one untimed pass with the fp32 CNN: the reference's fused pick is an fp32 result, so the grasp pixel is compared there
    where = list(range(len(gold)))
    r32 = np.frombuffer(eng.process_batch(lab_d, dep_d, cam, False, sync=False).cpu().numpy().tobytes(), dtype=N.FRAME_RESULT)
    check = {"golden_device_batch": check_golden(records, gold, where, compare_grasp=False),
             "golden_host_batch": check_golden(out_host, gold, where, compare_grasp=False),
             "golden_fp32_cnn": check_golden(r32, gold, where, compare_grasp=True)}
    check["golden_frames_identical"] = bool(all(check[k]["identical"] for k in
                                                ("golden_device_batch", "golden_host_batch", "golden_fp32_cnn")))
    ok = (records["ml_valid"] > 0) & (r32["ml_valid"] > 0)
    check["candidates_identical_to_fp32_path"] = bool(np.array_equal(records["cand_x"], r32["cand_x"]) and
                                                      np.array_equal(records["cand_y"], r32["cand_y"]))
    check["host_batch_identical_to_device_batch"] = records_equal(records, out_host)
    check["bf16_logit_max_abs_diff"] = float(np.abs(records["logit"][ok] - r32["logit"][ok]).max()) if ok.any() else 0.0
    check["fused_pick_agreement_bf16_vs_fp32"] = float((records["best_index"] == r32["best_index"]).mean())
    # (c) the record block the fusion kernel wrote for the all-gather equals the result structs
    exp = lgd.records_from_results(records, "cpu").numpy()
    check["gather_records_match_results"] = bool(np.array_equal(rec_dev, exp))

    cpu = None
    if want_cpu:
        r = cpu_arm(spec_name, args.cpu_steps, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    frames_per_step = B * cps * world
    ctx_gb = eng.context_bytes / 1e9
    eng.close()
    del lab_d, dep_d, rec_all
    torch.cuda.empty_cache()
    line = {
        "metric": METRIC, "value": frames_per_step / step_s, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong" if config == "cfg5" else "weak", "vs_baseline": None,
        "dtype": "f64/f32 scoring, u32 Q16 chamfer, " + ("bf16 CNN" if use_bf16 else "fp32 CNN"), "data": "synthetic",
        "config": {"workload": WORKLOADS[config], "frames_per_step": frames_per_step, "frames_per_gpu_per_call": B,
                   "calls_per_step_per_gpu": cps, "unique_frames_per_gpu": unique,
                   "golden_frames_in_batch": len(gold),
                   "l2": f"inputs ({B * P * 6 / 1e9:.2f} GB per call) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"frame-sharded x{world}; candidate records all-gathered once per timed region",
                   "lanes_per_gpu": args.lanes, "cnn": args.cnn, "picked": int((records["n_candidates"] > 0).sum()),
                   "context_gb": round(ctx_gb, 2), "cpu_affinity": f"{len(cpus)} CPUs of NUMA node(s) "
                   f"{hostmem.numa_node_of_cpus(cpus)}" if cpus else "unbound"},
        "e2e": {"value": frames_per_step * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_call * cps),
                "d2h_bytes_per_step": int(d2h_call * cps),
                "host_input_bytes_per_step": int(B * cps * P * 6),      # the caller's buffers: int16 labels + float32 depth
                "labels_on_the_link": "run-length encoded by host threads inside the call, expanded on the device (lossless)"
                if h2d_call < B * P * 6 else "raw",
                "h2d_gbs_per_gpu": round(h2d_call * cps * steps / e2e_s / 1e9, 1),
                "host_buffers_pinned": bool(pinned_ok),
                # all ranks copying the same inputs at once and doing nothing else, measured in this run: the ceiling of e2e
                "h2d_ceiling_gbs": round(h2d_ceiling_gbs, 1),
                "h2d_sum_of_links_gbs": round(h2d_links_gbs, 1),      # > ceiling when the links are not equally fast
                # the ceiling in frames: the link rate over the bytes a frame really puts on the link
                "frames_per_s_at_h2d_ceiling": round(h2d_ceiling_gbs * 1e9 / (h2d_call / B), 0),
                "frac_of_h2d_ceiling": round(frames_per_step * steps / e2e_s / (h2d_ceiling_gbs * 1e9 / (h2d_call / B)), 3)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "consistency": check,
    }
    return line


# ------------------------------------------------------------------------------------------------------
# cfg1: single-frame latency through the drop-in classes
# ------------------------------------------------------------------------------------------------------
def run_cfg1(args, D, iters=None):
    torch = D.torch
    if D.dev is None:
        D.init_cuda()
    from leafgrasp_b200 import (GraspPointCNN, GraspPointSelector, ImageProcessor, OptimalLeafSelector, synth)
    from leafgrasp_b200 import _native as N
    iters = iters or max(args.steps, 20)
    spec = synth.CFG1
    P = synth.projection_matrix(spec)
    gold = golden_frames("CFG1")[:1]
    seed, idx = (GOLDEN_SEED, gold[0][0]) if gold else (CONFIG_SEED, 0)
    lab, dep = synth.make_frame(spec, seed, idx)
    dev = D.dev
    lib = N.lib()
    scorer = OptimalLeafSelector(dev)
    scorer.set_camera_params(P)
    sel = GraspPointSelector(dev)
    sel.set_camera_params(P)
    net = GraspPointCNN(in_channels=9)
    net.load_state_dict(synth.seeded_state_dict(CNN_SEED))
    net.eval()
    sel.ml_predictor = net
    ip = ImageProcessor(spec.height, spec.width, 21, 5)
    lab_h, dep_h = torch.from_numpy(lab).pin_memory(), torch.from_numpy(dep).pin_memory()
    lab_d, dep_d = lab_h.to(dev), dep_h.to(dev)

    def node_step(mt, dt):      # leaf_grasp_node_v3.py:110-119
        leaf = scorer.select_optimal_leaf(mt, dt)
        return leaf, sel.select_grasp_point(mt == leaf, dt, ip)

    out = {}
    for bf16 in (False, True):
        sel.use_bf16_cnn = bf16
        for _ in range(max(3, args.warmup)):
            node_step(lab_d, dep_d)
        l0 = lib.lg_launch_count()
        lat = []
        for _ in range(iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            leaf, (g2, g3, pre) = node_step(lab_d, dep_d)
            torch.cuda.synchronize()
            lat.append(time.perf_counter() - t0)
        launches = int(lib.lg_launch_count() - l0)
        lat_h = []
        for _ in range(iters):
            t0 = time.perf_counter()
            mt, dt = lab_h.to(dev, non_blocking=True), dep_h.to(dev, non_blocking=True)
            leaf_h, (g2h, _, _) = node_step(mt, dt)
            torch.cuda.synchronize()
            lat_h.append(time.perf_counter() - t0)
        r = sel.last_result
        out["bf16" if bf16 else "fp32"] = dict(lat=statistics.median(lat), lat_host=statistics.median(lat_h), leaf=leaf,
                                               g2=g2, rec=r.copy() if r is not None else None, launches=launches)
    check = None
    if gold:
        g = gold[0][1]
        f32, b16 = out["fp32"], out["bf16"]
        npos = int(g["n_positive"])
        cands = lambda r: np.stack([r["cand_x"], r["cand_y"]], axis=1)[:npos]
        check = {"golden_frames_in_batch": 1,
                 "leaf_id_identical": bool(f32["leaf"] == int(g["leaf_id"]) and b16["leaf"] == int(g["leaf_id"])),
                 "positive_candidates_identical": bool(np.array_equal(cands(f32["rec"]), g["candidates"][:npos]) and
                                                       np.array_equal(cands(b16["rec"]), g["candidates"][:npos])),
                 "grasp_identical_fp32_cnn": bool(tuple(f32["g2"]) == tuple(int(v) for v in g["grasp_2d"])),
                 "grasp_identical_bf16_cnn": bool(tuple(b16["g2"]) == tuple(int(v) for v in g["grasp_2d"]))}
        check["golden_frames_identical"] = bool(check["leaf_id_identical"] and check["positive_candidates_identical"] and
                                                check["grasp_identical_fp32_cnn"])
    best = out[args.cnn]
    line = {"metric": METRIC, "value": 1.0 / best["lat"], "unit": UNIT, "n_gpus": 1, "steps": iters, "warmup": max(3, args.warmup),
            "ms_per_step": best["lat"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64/f32 scoring, u32 Q16 chamfer, " + args.cnn + " CNN", "data": "synthetic",
            "config": {"workload": WORKLOADS["cfg1"], "frames_per_step": 1, "timing": "median wall time per frame, synchronised",
                       "l2": "single 9.3 MB frame: L2-resident after the first touch, as in the live node",
                       "latency_ms": {k: round(v["lat"] * 1e3, 3) for k, v in out.items()},
                       "latency_ms_host_tensors": {k: round(v["lat_host"] * 1e3, 3) for k, v in out.items()}},
            "e2e": {"value": 1.0 / best["lat_host"], "unit": UNIT, "h2d_bytes_per_step": int(spec.height * spec.width * 6),
                    "d2h_bytes_per_step": int(2 * N.FRAME_RESULT.itemsize)},
            "gpu_launches": best["launches"], "consistency": check}
    return line


# ------------------------------------------------------------------------------------------------------
# cfg4: GraspPointCNN only
# ------------------------------------------------------------------------------------------------------
def run_cfg4(args, D, n_patches=65536, steps=None):
    torch = D.torch
    if D.dev is None:
        D.init_cuda()
    from leafgrasp_b200 import GraspEngine, pack_weights, synth
    from leafgrasp_b200 import _native as N
    dev = D.dev
    lib = N.lib()
    steps = steps or args.steps
    pk = peaks()
    eng = GraspEngine(max(1, n_patches // N.TOP_K + 1), 64, 64, 2, device=dev)       # activation scratch for all patches at once
    eng.set_cnn_weights(pack_weights(synth.seeded_state_dict(CNN_SEED)))
    g = torch.Generator().manual_seed(4)
    gold = np.load(os.path.join(GOLD, "cnn_patches.npz"))            # 16 patches with the reference module's logits
    n_gold = gold["x"].shape[0]
    x_h = torch.empty(n_patches, 9, 32, 32, dtype=torch.float32, pin_memory=True)
    chunk = 4096
    for lo in range(0, n_patches, chunk):
        hi = min(n_patches, lo + chunk)
        x_h[lo:hi] = torch.rand(hi - lo, 9, 32, 32, generator=g)
    x_h[:, 1] = (x_h[:, 1] > 0.5).float()
    x_h[:n_gold] = torch.from_numpy(gold["x"])
    x_d = x_h.to(dev)
    times, logits = {}, {}
    launches = 0
    for mode in ("bf16", "fp32"):
        bf = mode == "bf16"
        for _ in range(max(3, args.warmup) if bf else 1):
            y = eng.cnn_forward(x_d, bf)
        torch.cuda.synchronize()
        n_it = steps if bf else max(1, min(steps, 2))
        l0 = lib.lg_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_it):
            y = eng.cnn_forward(x_d, bf)
        e1.record()
        torch.cuda.synchronize()
        times[mode] = e0.elapsed_time(e1) / n_it
        logits[mode] = y.cpu().numpy()
        if bf:
            launches = int(lib.lg_launch_count() - l0)
    # end to end: pinned host patches in, host logits out
    y_h = torch.empty(n_patches, dtype=torch.float32, pin_memory=True)
    x_stage = torch.empty_like(x_d)
    for _ in range(2):
        x_stage.copy_(x_h, non_blocking=True)
        y_h.copy_(eng.cnn_forward(x_stage, True), non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        x_stage.copy_(x_h, non_blocking=True)
        y_h.copy_(eng.cnn_forward(x_stage, True), non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    tf = n_patches * CNN_FLOP / (times["bf16"] * 1e-3) / 1e12
    check = {"golden_patches_in_batch": int(n_gold),
             "bf16_vs_reference_logits_max_abs": float(np.abs(logits["bf16"][:n_gold] - gold["logits"]).max()),
             "fp32_vs_reference_logits_max_abs": float(np.abs(logits["fp32"][:n_gold] - gold["logits"]).max()),
             "bf16_vs_fp32_max_abs": float(np.abs(logits["bf16"] - logits["fp32"]).max())}
    check["golden_frames_identical"] = bool(check["bf16_vs_reference_logits_max_abs"] <= 1e-2 and
                                            check["fp32_vs_reference_logits_max_abs"] <= 1e-4)
    eng.close()
    del x_d, x_stage
    torch.cuda.empty_cache()
    return {"metric": "cnn_patches_per_sec", "value": n_patches / (times["bf16"] * 1e-3), "unit": "patches/s", "n_gpus": 1,
            "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": times["bf16"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (tcgen05, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOADS["cfg4"], "patches": n_patches,
                       "l2": f"{n_patches * 9 * 1024 * 4 / 1e9:.1f} GB of patches per step exceed the L2",
                       "fp32_cuda_core_ms_per_step": round(times["fp32"], 2),
                       "tensor_core_speedup_over_cuda_cores": round(times["fp32"] / times["bf16"], 2)},
            "e2e": {"value": n_patches / e2e_s, "unit": "patches/s", "h2d_bytes_per_step": int(n_patches * 9 * 1024 * 4),
                    "d2h_bytes_per_step": int(n_patches * 4)},
            "gpu_launches": launches,
            "roofline": {"kernel": "conv3x3_umma_kernel x6 + tail", "bound": "tensor", "achieved": tf, "peak": pk["tf_sustained"],
                         "unit": "TFLOP/s", "frac": tf / pk["tf_sustained"], "traffic": None,
                         "peak_source": pk["source"] + " (sustained: the step lasts >20 ms)",
                         "frac_of_burst_peak": tf / pk["tf_burst"]},
            "consistency": check}


def brief(line):
    """The part of a config's line kept under the default run's `extra`."""
    keep = ("metric", "value", "unit", "ms_per_step", "steps", "config", "e2e", "consistency", "gpu_launches")
    out = {k: line[k] for k in keep if k in line}
    if "roofline" in line and line["roofline"]:
        r = line["roofline"]
        out["roofline"] = {k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "stage_ms_serial",
                                             "stage_frac_of_hbm_peak", "cnn_tflops", "frac_of_burst_peak") if k in r}
    return out


def run_ours(args):
    D = Dist()
    cfg = args.config
    if cfg in ("cfg1", "cfg4"):
        if D.rank != 0:        # single-GPU configurations: replicas would only repeat rank 0
            return
        line = run_cfg1(args, D) if cfg == "cfg1" else run_cfg4(args, D)
        line["n_gpus"] = 1
        if not args.no_cpu:
            if cfg == "cfg1":
                r = cpu_arm("CFG1", max(1, args.cpu_steps), 1, workers=1)
            else:
                r = cpu_cnn_arm(max(1, args.cpu_steps), 1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
        return
    if cfg == "cfg3":
        line = run_frames(args, D, cfg, "CFG3", args.frames or 128, 1, args.unique or 16, D.world == 1 and not args.no_cpu, 1)
    elif cfg == "cfg5":
        B = args.frames or 256
        total = 8192
        if total % (B * D.world):
            raise SystemExit(f"cfg5: 8192 frames do not split into calls of {B} frames on {D.world} GPUs")
        line = run_frames(args, D, cfg, "CFG2", B, total // (B * D.world), args.unique or 256, D.world == 1 and not args.no_cpu, 10)
    else:
        line = run_frames(args, D, cfg, "CFG2", args.frames or 256, 1, args.unique or 256, D.world == 1 and not args.no_cpu, 10)
    if D.rank == 0 and cfg == "cfg2" and D.world == 1 and not args.no_extra:
        # the other single-GPU configurations, short versions (their own lines: bench.py --config cfgN)
        extra = {}
        small = argparse.Namespace(**vars(args))
        small.steps, small.warmup, small.lanes = 3, 3, 1
        for name, fn in (("cfg1", lambda: run_cfg1(small, D, iters=20)),
                         ("cfg3", lambda: run_frames(small, D, "cfg3", "CFG3", 64, 1, 8, False, 1)),
                         ("cfg4", lambda: run_cfg4(small, D, steps=3))):
            try:
                extra[name] = brief(fn())
            except Exception as e:  # noqa: BLE001 - an extra must not take the headline line down
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
        line["extra"] = extra
    if D.rank == 0:
        print(json.dumps(line))
    D.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per call (default 256; cfg3: 128 = 69 GB of context)")
    ap.add_argument("--unique", type=int, default=0, help="distinct synthetic frames generated per rank (default: all of them)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cnn", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--lanes", type=int, default=1,
                    help="parts a GPU's batch is processed in, side by side on streams (the per-stage times of the timed "
                         "region then include the other part's kernels)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="default run only: skip the short cfg1 / cfg3 / cfg4 measurements")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the process to the CPUs next to its GPU")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
