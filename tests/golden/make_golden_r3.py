"""Third set of golden vectors, produced like the first two by RUNNING THE REFERENCE ITSELF (make_golden.py: same harness,
same environment, same seeds).  Run by hand in the build container:

    python tests/golden/make_golden_r3.py

Adds, without touching the earlier sets, frames of the shapes that had one reference answer so far:
  * frame_cfg1_{1..4}.npz    four more 1440x1080 / 10-leaf frames (BASELINE config 1, the live node's shape);
  * frame_small_{4..7}.npz   four more 480x360 frames;
  * frame_cfg3_1.npz         a second 3840x2160 / 100-leaf frame (BASELINE config 3).
golden_meta_r3.json lists them.  They pin the oracle on the CPU (tests/test_oracle_golden.py); the GPU tests of the
round reach them through the oracle.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import make_golden as G  # noqa: E402  (sets up sys.path, imports harness + synth + oracle)


def main():
    classes = G.ref_harness.load()
    sd = G.O.seeded_state_dict(G.CNN_SEED)
    meta = {"config_seed": G.CONFIG_SEED, "cnn_seed": G.CNN_SEED, "frames": []}
    plan = [("CFG1", i, False) for i in range(1, 5)] + [("SMALL", i, False) for i in range(4, 8)] + [("CFG3", 1, False)]
    for name, idx, full in plan:
        spec = getattr(G.synth, name)
        out = G.run_frame(spec, idx, classes, sd, full)
        fn = f"frame_{name.lower()}_{idx}.npz"
        np.savez_compressed(os.path.join(HERE, fn), **out)
        meta["frames"].append({"spec": name, "index": idx, "file": fn, "leaf_id": out["leaf_id"]})
        print(fn, "leaf", out["leaf_id"], "n_pos", out.get("n_positive"), "n_ml", out.get("n_ml"), flush=True)
    with open(os.path.join(HERE, "golden_meta_r3.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
