"""Small driver for ncu captures of the whole step: lg_process_batch on a batch of cfg2 frames (bf16 CNN), three times.
python tools/step_profile.py [frames] [unique frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
u = int(sys.argv[2]) if len(sys.argv) > 2 else 16
spec = synth.CFG2
lab, dep = synth.make_batch(spec, 11, 0, min(n, u))
reps = (n + lab.shape[0] - 1) // lab.shape[0]
lab = torch.from_numpy(np.tile(lab, (reps, 1, 1))[:n]).cuda()
dep = torch.from_numpy(np.tile(dep, (reps, 1, 1))[:n]).cuda()
eng = GraspEngine(n, spec.height, spec.width, 128)
eng.set_cnn_weights(pack_weights(synth.seeded_state_dict(1234)))
cam = camera_from_projection(synth.projection_matrix(spec))
for _ in range(3):
    res = eng.process_batch(lab, dep, cam, True)
torch.cuda.synchronize()
print("leaf ids", [int(r["leaf_id"]) for r in res[:8]])
