"""Serial per-stage times (overlap off) of lg_process_batch on the metric's workload; LG_LIB=<path to a liblgb200 build> selects
the library (experiments with -D switches), default the in-tree one.  python tools/stage_times.py [frames]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import _native as N
if os.environ.get("LG_LIB"):
    N.LIB_PATH = os.path.abspath(os.environ["LG_LIB"])
from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spec, U = synth.CFG2, 32
lab, dep = synth.make_batch(spec, 7, 0, U)
lab = np.concatenate([lab] * (B // U)); dep = np.concatenate([dep] * (B // U))
lt, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda()
eng = GraspEngine(B, spec.height, spec.width, 128)
eng.set_cnn_weights(pack_weights(O.seeded_state_dict(1234)))
if os.environ.get('LG_PATCH_EXPORT', '0') != '1':
    eng.set_patch_export(False)
cam = camera_from_projection(synth.projection_matrix(spec))
lib = N.lib()
names = ["leaf_stats", "scatter", "median", "edt_columns", "edt_rows", "select", "chamfer", "orientation", "score_maps",
         "candidates", "patches", "cnn", "fuse"]
for overlap in (False, True):
    eng.set_overlap(overlap)
    lib.lg_set_profiling(eng._ctx, 1)
    acc = np.zeros(14); tot = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(8):
        e0.record(); eng.process_batch(lt, dt, cam, True, sync=False); e1.record()
        buf = (C.c_float * 14)(); lib.lg_stage_times(eng._ctx, buf, 14)
        torch.cuda.synchronize()
        if it >= 3:
            acc += np.array(list(buf)); tot.append(e0.elapsed_time(e1))
    acc /= 5
    print(("overlap " if overlap else "serial  ") + f"step {np.mean(tot):.3f} ms | " + " ".join(f"{n}={acc[i + 1]:.3f}" for i, n in enumerate(names)), flush=True)
