"""Seeded inputs of the smooth_depth golden cases, shared by the generator (make_golden_r2.py) and the tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG_SEED = 7


def smooth_input(name, h, w):
    """Uniform noise of the given size, or (name == "leaf") the masked depth of SMALL frame 0, leaf 1."""
    if name == "leaf":
        from leafgrasp_b200 import synth
        lab, dep = synth.make_frame(synth.SMALL, CONFIG_SEED, 0)
        return (dep * (lab == 1)).astype(np.float32)
    rng = np.random.default_rng(h * 1000 + w)
    return rng.uniform(0.2, 0.9, size=(h, w)).astype(np.float32)
