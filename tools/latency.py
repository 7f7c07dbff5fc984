"""Single-frame latency of the drop-in call sequence (leaf_grasp_node_v3.py:110-119) and of lg_process_batch at B = 1."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import (GraspEngine, GraspPointCNN, GraspPointSelector, ImageProcessor, OptimalLeafSelector,
                            camera_from_projection, pack_weights, synth)

for spec_name in ("CFG1", "CFG2"):
    spec = getattr(synth, spec_name)
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_frame(spec, 11, 0)
    sd = O.seeded_state_dict(1234)
    dev = torch.device("cuda")
    mask_t, depth_t = torch.from_numpy(lab).to(dev), torch.from_numpy(dep).to(dev)
    eng = GraspEngine(1, spec.height, spec.width, 128)
    eng.set_cnn_weights(pack_weights(sd))
    cam = camera_from_projection(P)
    for _ in range(5):
        eng.process_batch(mask_t[None], depth_t[None], cam, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        eng.process_batch(mask_t[None], depth_t[None], cam, True)
    torch.cuda.synchronize()
    t_engine = (time.perf_counter() - t0) / 50
    scorer = OptimalLeafSelector(dev); scorer.set_camera_params(P)
    sel = GraspPointSelector(dev); sel.set_camera_params(P)
    net = GraspPointCNN(in_channels=9); net.load_state_dict(sd); net.eval(); sel.ml_predictor = net
    ip = ImageProcessor(spec.height, spec.width, 21, 5)
    def node_step():
        leaf = scorer.select_optimal_leaf(mask_t, depth_t)
        return sel.select_grasp_point(mask_t == leaf, depth_t, ip)
    for _ in range(5):
        node_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30):
        out = node_step()
    torch.cuda.synchronize()
    t_node = (time.perf_counter() - t0) / 30
    print(f"{spec_name}: lg_process_batch B=1 {t_engine * 1e3:.2f} ms/frame; drop-in classes (select_optimal_leaf + select_grasp_point) "
          f"{t_node * 1e3:.2f} ms/frame; grasp {out[0]}", flush=True)
    eng.close()
