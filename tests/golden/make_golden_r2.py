"""Second set of golden vectors, again produced by RUNNING THE REFERENCE ITSELF (see make_golden.py; same harness,
same environment).  Run by hand in the build container:

    python tests/golden/make_golden_r2.py

Adds, without touching the first set:
  * frame_cfg2_{2..9}.npz   eight more 1440x1080 / 30-leaf frames (the benchmark's workload);
  * frame_cfg3_0.npz        one 3840x2160 / 100-leaf frame (BASELINE config 3);
  * smooth_depth.npz        ImageProcessor.smooth_depth (image_processor.py:56-64) on three seeded images.
golden_meta_r2.json lists them.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import make_golden as G  # noqa: E402  (sets up sys.path, imports harness + synth + oracle)

SMOOTH_CASES = [("small", 32, 32), ("odd", 37, 53), ("tiny", 3, 5), ("leaf", 360, 480)]


from make_golden_r2_inputs import smooth_input  # noqa: E402  (seeded inputs, shared with the tests)


def main():
    classes = G.ref_harness.load()
    sd = G.O.seeded_state_dict(G.CNN_SEED)
    meta = {"config_seed": G.CONFIG_SEED, "cnn_seed": G.CNN_SEED, "frames": [], "smooth_depth": []}
    plan = [("CFG2", i, False) for i in range(2, 10)] + [("CFG3", 0, False)]
    for name, idx, full in plan:
        spec = getattr(G.synth, name)
        out = G.run_frame(spec, idx, classes, sd, full)
        fn = f"frame_{name.lower()}_{idx}.npz"
        np.savez_compressed(os.path.join(HERE, fn), **out)
        meta["frames"].append({"spec": name, "index": idx, "file": fn, "leaf_id": out["leaf_id"]})
        print(fn, "leaf", out["leaf_id"], "n_pos", out.get("n_positive"), "n_ml", out.get("n_ml"), flush=True)
    _, _, IP, _ = classes
    outs = {}
    for name, h, w in SMOOTH_CASES:
        ip = IP(h, w, 21, 5)
        x = smooth_input(name, h, w)
        y = ip.smooth_depth(torch.from_numpy(x), torch.device("cpu"))
        outs[name] = y.numpy().astype(np.float32)
        meta["smooth_depth"].append({"name": name, "height": h, "width": w})
    np.savez_compressed(os.path.join(HERE, "smooth_depth.npz"), **outs)
    with open(os.path.join(HERE, "golden_meta_r2.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
