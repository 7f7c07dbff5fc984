"""Small driver for ncu captures of the stage-1 kernels: lg_select_leaf on a batch of cfg2 frames, three times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from leafgrasp_b200 import GraspEngine, camera_from_projection, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
spec = getattr(synth, sys.argv[2]) if len(sys.argv) > 2 else synth.CFG2
lab, dep = synth.make_batch(spec, 11, 0, min(n, 16))
reps = (n + lab.shape[0] - 1) // lab.shape[0]
lab = torch.from_numpy(np.tile(lab, (reps, 1, 1))[:n]).cuda()
dep = torch.from_numpy(np.tile(dep, (reps, 1, 1))[:n]).cuda()
eng = GraspEngine(n, spec.height, spec.width, 128)
cam = camera_from_projection(synth.projection_matrix(spec))
for _ in range(3):
    ids, _ = eng.select_leaf(lab, dep, cam)
torch.cuda.synchronize()
print("leaf ids", ids[:8].tolist())
