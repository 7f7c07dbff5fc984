"""Experiment: does running two half-batches on two streams (two contexts) beat one full batch?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth

spec = synth.CFG2
B = 256
cam = camera_from_projection(synth.projection_matrix(spec))
lab_u, dep_u = synth.make_batch(spec, 11, 0, 16)
lab = torch.from_numpy(np.tile(lab_u, (B // 16, 1, 1))).cuda()
dep = torch.from_numpy(np.tile(dep_u, (B // 16, 1, 1))).cuda()
blob = pack_weights(O.seeded_state_dict(1234))

def timeit(fn, iters=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for parts in (1, 2, 4):
    n = B // parts
    engs = [GraspEngine(n, spec.height, spec.width, 128) for _ in range(parts)]
    for e in engs: e.set_cnn_weights(blob)
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def run():
        cur = torch.cuda.current_stream()
        for k, (e, s) in enumerate(zip(engs, streams)):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                e.process_batch(lab[k * n:(k + 1) * n], dep[k * n:(k + 1) * n], cam, True, sync=False)
        for s in streams: cur.wait_stream(s)
    ms = timeit(run)
    print(f"parts={parts}: {ms:.3f} ms per {B} frames -> {B / ms * 1e3:.0f} frames/s", flush=True)
    for e in engs: e.close()
    del engs
