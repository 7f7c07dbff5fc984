"""BASELINE config[2] as a timing run (not a bench.py line): 3840x2160 frames with 100 overlapping leaves."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import ctypes as C
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
spec = synth.CFG3
cam = camera_from_projection(synth.projection_matrix(spec))
lab_u, dep_u = synth.make_batch(spec, 11, 0, 4)
lab = torch.from_numpy(np.tile(lab_u, (B // 4, 1, 1))).cuda()
dep = torch.from_numpy(np.tile(dep_u, (B // 4, 1, 1))).cuda()
eng = GraspEngine(B, spec.height, spec.width, 128)
eng.set_cnn_weights(pack_weights(O.seeded_state_dict(1234)))
for _ in range(3):
    eng.process_batch(lab, dep, cam, True, sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.process_batch(lab, dep, cam, True, sync=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
eng.set_overlap(False); eng.lib.lg_set_profiling(eng._ctx, 1)
res = eng.process_batch(lab, dep, cam, True)
buf = (C.c_float * 14)(); eng.lib.lg_stage_times(eng._ctx, buf, 14)
names = ["", "leaf_stats", "scatter", "median", "edt_columns", "edt_rows", "select", "chamfer", "orientation", "score_maps",
         "candidates", "patches", "cnn", "fuse"]
print(f"cfg3: {B} frames 3840x2160, 100 leaves: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} frames/s "
      f"({B / ms * 1e3 * spec.height * spec.width / 1e9:.1f} Gpx/s); picked {int((res['n_candidates'] > 0).sum())}/{B}")
print({names[i]: round(buf[i], 3) for i in range(1, 14)})
print("context GB", eng.context_bytes / 1e9)
