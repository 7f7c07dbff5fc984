"""Strict-oracle answers for many frames at once: one spawned worker per host core (the oracle takes ~1 s per
1440x1080 frame and ~11 s per 4K frame single-threaded).  Test infrastructure only."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _init():
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cv2
    import torch
    cv2.setNumThreads(1)
    torch.set_num_threads(1)


def _one(job):
    import numpy as np
    import leafgrasp_oracle as O
    from leafgrasp_b200 import synth
    spec_name, seed, idx, cnn_seed = job
    spec = getattr(synth, spec_name)
    lab, dep = synth.make_frame(spec, seed, idx)
    o = O.process_frame(lab, dep, synth.projection_matrix(spec), O.seeded_state_dict(cnn_seed), arith="strict")
    if o["leaf_id"] is None:
        return {"leaf_id": -1}
    d = o["debug"]
    return {"leaf_id": int(o["leaf_id"]), "picks": [tuple(p) for p in d["picks"]], "trad_at": [float(t) for t in d["trad_at"]],
            "logits": [None if v is None else float(v) for v in d["logits"]], "grasp": tuple(int(v) for v in o["grasp"][0]),
            "grasp_3d": tuple(float(v) for v in o["grasp"][1]), "pre_grasp": tuple(float(v) for v in o["grasp"][2]),
            "n_positive": int(sum(1 for t, (x, y) in zip(d["trad_at"], d["picks"]) if d["valid"][y, x] and t > 0))}


def strict_answers(spec_name, seed, indices, cnn_seed, workers=None):
    jobs = [(spec_name, seed, int(i), cnn_seed) for i in indices]
    workers = workers or max(1, min(len(jobs), os.cpu_count() or 1, 32))
    with mp.get_context("spawn").Pool(workers, initializer=_init) as pool:
        return pool.map(_one, jobs, chunksize=1)
