#!/usr/bin/env python
"""Benchmark of the grasp-selection hot path (BASELINE.json metric: frames/s/GPU at 1440x1080, 30 leaves).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference] [--cnn bf16|fp32]

One "step" = one pass of the whole path (leaf selection -> score maps -> top-20 candidates -> patches ->
GraspPointCNN -> fusion) over one batch of B synthetic frames per GPU (BASELINE config[1]: B = 256 frames,
30 leaves each).  Prints ONE JSON line (rank 0):
  value     whole-job frames/s with the batch resident in HBM, timed on the device with CUDA events,
            barrier + synchronize on both sides, max over ranks;
  e2e       the same through lg_process_batch_host: pinned host buffers in, host records out, copies timed;
  roofline  the dominant kernel of the step (per-stage CUDA events inside the library): algorithmic bytes
            per launch / its device time, against MEASURED_PEAKS.json;
  cpu_baseline  the oracle port of the reference's CPU path timed on this box's host cores (bounded sample).
--impl reference times that CPU path alone (the reference is pure Python and /root/reference does not travel
to the GPU box, so the arm runs the oracle port: kind "port").
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
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "grasp_selection_frames_per_sec"
UNIT = "frames/s"
WORKLOAD = "cfg2: 1440x1080 frames, 30 leaves, full path (leaf selection + score maps + top-20 + patches + GraspPointCNN + fusion)"
CONFIG_SEED = 11
CNN_SEED = 1234
STAGES = ["", "leaf_stats", "scatter", "median", "edt_columns", "edt_rows", "select", "chamfer", "orientation",
          "score_maps", "candidates", "patches", "cnn", "fuse"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
# CPU arm (oracle port of the reference's path)
# ------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(spec_name, seed):
    import cv2
    import torch
    import leafgrasp_oracle as O
    from leafgrasp_b200 import synth
    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    _W["O"], _W["synth"] = O, synth
    _W["spec"] = getattr(synth, spec_name)
    _W["P"] = synth.projection_matrix(_W["spec"])
    _W["sd"] = O.seeded_state_dict(CNN_SEED)
    # one seeded frame per worker process, generated before any timing
    _W["frame"] = synth.make_frame(_W["spec"], seed, os.getpid() % 4096)


def _cpu_frame(idx):
    lab, dep = _W["frame"]
    t = time.perf_counter()
    _W["O"].process_frame(lab, dep, _W["P"], _W["sd"], arith="reference")
    return time.perf_counter() - t


def cpu_arm(steps, warmup, workers=None):
    """Each step: `workers` frames, one per worker process (1 thread each) -> frames/s over all host cores."""
    import multiprocessing as mp
    workers = workers or max(1, min(os.cpu_count() or 1, 32))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers, initializer=_cpu_init, initargs=("CFG2", CONFIG_SEED)) as pool:
        frames = list(range(workers))
        for _ in range(max(1, warmup)):
            pool.map(_cpu_frame, frames, chunksize=1)
        t0 = time.perf_counter()
        per = []
        for _ in range(steps):
            per += pool.map(_cpu_frame, frames, chunksize=1)
        wall = time.perf_counter() - t0
    return {"value": workers * steps / wall, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": f"{steps} steps x {workers} cfg2 frames, one oracle process per core (1 thread each); "
                      f"mean {statistics.mean(per):.2f} s/frame/core",
            "ms_per_step": wall / steps * 1e3, "frames_per_step": workers}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 (NumPy, OpenCV, torch CPU)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": r["frames_per_step"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import leafgrasp_oracle as O          # only for the seeded state_dict and the cpu_baseline leg
    from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth
    from leafgrasp_b200 import _native as N
    from leafgrasp_b200 import dist as lgd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = synth.CFG2
    B = args.frames
    H, W = spec.height, spec.width
    cam = camera_from_projection(synth.projection_matrix(spec))
    use_bf16 = args.cnn == "bf16"

    # synthetic batch: `unique` distinct seeded frames tiled to B (the path has no cross-frame state)
    unique = min(B, args.unique)
    lab_u, dep_u = synth.make_batch(spec, CONFIG_SEED, rank * unique, unique)
    reps = (B + unique - 1) // unique
    lab_h = torch.from_numpy(np.tile(lab_u, (reps, 1, 1))[:B]).pin_memory()
    dep_h = torch.from_numpy(np.tile(dep_u, (reps, 1, 1))[:B]).pin_memory()
    lab_d, dep_d = lab_h.to(dev), dep_h.to(dev)

    eng = GraspEngine(B, H, W, 128, device=dev, lanes=args.lanes)
    n_prof = eng.lane_split(B)[0][1]          # frames the profiled (main) context handles in the timed region
    eng.set_cnn_weights(pack_weights(O.seeded_state_dict(CNN_SEED)))
    lib = N.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        res = eng.process_batch(lab_d, dep_d, cam, use_bf16, sync=False)
        if world > 1:       # aggregation of the candidate records, as north_star specifies (stays on the device)
            lgd.gather_candidate_records(lgd.records_from_result_buffer(res, B), B * world)
        return res

    for _ in range(args.warmup):
        step_device()
    # ---- timed region 1: inputs resident in HBM -----------------------------------------------------
    lib.lg_set_profiling(eng._ctx, 1)
    stage_ms = np.zeros(len(STAGES))
    sampler = ClockSampler(local)
    launches0 = lib.lg_launch_count()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_device()
        buf = (C.c_float * 14)()
        lib.lg_stage_times(eng._ctx, buf, 14)      # waits for the step's last event: steps are serial anyway
        stage_ms += np.array(list(buf))
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = int(lib.lg_launch_count() - launches0)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    lib.lg_set_profiling(eng._ctx, 0)
    records = np.frombuffer(res.cpu().numpy().tobytes(), dtype=N.FRAME_RESULT)

    # ---- timed region 2: end to end through the host-buffer entry point ----------------------------
    for _ in range(2):
        eng.process_batch_host(lab_h, dep_h, cam, use_bf16)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.process_batch_host(lab_h, dep_h, cam, use_bf16)
        if world > 1:
            lgd.gather_candidate_records(lgd.records_from_results(out, dev), B * world)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage -------------------------------------------------------------
    # stage_ms: CUDA events recorded inside the library during the timed region (stages on the library's side stream
    # overlap the others).  The dominant kernel is picked from one extra, untimed pass with the overlap switched off,
    # i.e. by the kernels' own durations; `achieved` uses its duration inside the timed region.
    hbm_peak, tf_peak, which = peaks()
    stage_ms /= args.steps
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
    P = H * W
    reg = records["region"].astype(np.int64)
    bbox_px = float(np.mean(np.maximum(reg[:, 2] - reg[:, 0], 0) * np.maximum(reg[:, 3] - reg[:, 1], 0)))
    rect_px = float(np.mean(np.maximum(np.minimum(reg[:, 2] + 16, W) - np.maximum(reg[:, 0] - 16, 0), 0) *
                            np.maximum(np.minimum(reg[:, 3] + 16, H) - np.maximum(reg[:, 1] - 16, 0), 0)))
    leaf_px = float(np.mean((lab_u > 0).sum(axis=(1, 2))))
    n_patches = float(np.mean(records["ml_valid"].sum(axis=1)))
    alg_bytes = {   # per frame, compulsory traffic: inputs once + outputs once
        # leaf_stats also carries the column pass of the union distance transform (writes 2 B/px of column distances)
        "leaf_stats": 8 * P, "scatter": 6 * P + 4 * leaf_px, "median": 4 * leaf_px, "edt_columns": 0,
        "edt_rows": 2 * P, "select": 0, "chamfer": 2 * P + 6 * bbox_px, "orientation": 2 * bbox_px,
        "score_maps": (2 + 4 + 4 + 45) * rect_px, "candidates": 12 * 20000, "patches": 20 * 9 * 1024 * 8, "fuse": 760}
    top = int(np.argmax(serial_ms))
    name = STAGES[top]
    traffic = None      # DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), same batch only
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(name)
        if ent and int(ent.get("frames", 0)) > 0:   # captured at ent["frames"] frames per launch; traffic is linear in frames
            traffic = float(ent["dram_bytes_per_launch"]) * n_prof / int(ent["frames"])
    except Exception:  # noqa: BLE001
        pass
    if name == "cnn":
        flops = 312.83e6 * n_patches * n_prof
        ach = flops / (stage_ms[top] * 1e-3) / 1e12
        roof = {"kernel": "cnn", "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": ach / tf_peak, "traffic": traffic, "peak_source": which}
    else:
        ach = alg_bytes[name] * n_prof / (stage_ms[top] * 1e-3) / 1e9
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": traffic, "peak_source": which}
    roof["stage_ms"] = {STAGES[i]: round(float(stage_ms[i]), 4) for i in range(1, len(STAGES))}
    roof["stage_ms_frames"] = n_prof      # stage_ms: the main context's share of the batch (lanes run side by side)
    roof["stage_ms_serial"] = {STAGES[i]: round(float(serial_ms[i]), 4) for i in range(1, len(STAGES))}
    roof["stage_gbs"] = {k: round(alg_bytes[k] * B / (serial_ms[STAGES.index(k)] * 1e-3) / 1e9, 1)
                         for k in alg_bytes if serial_ms[STAGES.index(k)] > 0 and alg_bytes[k] > 0}
    roof["cnn_tflops"] = round(312.83e6 * n_patches * B / (serial_ms[STAGES.index("cnn")] * 1e-3) / 1e12, 1) \
        if serial_ms[STAGES.index("cnn")] > 0 else None
    roof["whole_step_gbs"] = round((45 * P) * B / (ms / args.steps * 1e-3) / 1e9, 1)   # 45 B/px, SURVEY.md 8d

    # ---- in-run consistency check of the batch just timed (untimed): tensor-core CNN against the fp32 CUDA path --
    check = None
    if use_bf16:
        r32 = np.frombuffer(eng.process_batch(lab_d, dep_d, cam, False, sync=False).cpu().numpy().tobytes(), dtype=N.FRAME_RESULT)
        ok = (records["ml_valid"] > 0) & (r32["ml_valid"] > 0)
        check = {"candidates_identical_to_fp32_path": bool(np.array_equal(records["cand_x"], r32["cand_x"]) and
                                                            np.array_equal(records["cand_y"], r32["cand_y"])),
                 "bf16_logit_max_abs_diff": float(np.abs(records["logit"][ok] - r32["logit"][ok]).max()) if ok.any() else 0.0,
                 "fused_pick_agreement": float((records["best_index"] == r32["best_index"]).mean())}

    cpu = None
    if world == 1 and not args.no_cpu:
        r = cpu_arm(args.cpu_steps, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    total_frames = B * world * args.steps
    value = total_frames / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64/f32 scoring, u32 Q16 chamfer, " + ("bf16 CNN" if use_bf16 else "fp32 CNN"), "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": B, "unique_frames": unique,
                   "l2": "inputs (2.4 GB per step at 256 frames) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"frame-sharded x{world}", "lanes_per_gpu": args.lanes, "cnn": args.cnn,
                   "picked": int((records["n_candidates"] > 0).sum())},
        "e2e": {"value": total_frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(B * P * 6),
                "d2h_bytes_per_step": int(B * N.FRAME_RESULT.itemsize)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "consistency": check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--unique", type=int, default=32, help="distinct synthetic frames generated per rank")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cnn", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--lanes", type=int, default=1,
                    help="parts a GPU's batch is processed in, side by side on streams (2: +7 %% frames/s, but the per-stage "
                         "times of the timed region then include the other part's kernels)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
