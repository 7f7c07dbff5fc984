"""Shared by the CPU and GPU collector tests: rebuild the inputs of the golden cases (tests/golden/collector.npz)."""
import hashlib
import os
import sys

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import collector_cases as CC  # noqa: E402

from leafgrasp_b200 import synth  # noqa: E402

CASES = [f"small{i}" for i in range(CC.SMALL_CASES)] + [f"crafted{i}" for i in range(CC.CRAFTED_CASES)]
_gold = None


def gold():
    global _gold
    if _gold is None:
        _gold = np.load(os.path.join(GOLD, "collector.npz"))
    return _gold


def case_inputs(name):
    """(mask u8, depth f32, P, golden dict of the case, labels int16 or None, leaf id or None)."""
    g = {k.split("/", 1)[1]: gold()[k] for k in gold().files if k.startswith(name + "/")}
    P = synth.projection_matrix(synth.SMALL)
    if name.startswith("small"):
        lab, dep = synth.make_frame(synth.SMALL, CC.CONFIG_SEED, int(name[5:]))
        leaf = int(g["leaf_id"])
        mask = (lab == leaf).astype(np.uint8)
    else:
        mask, dep = CC.crafted(int(name[7:]))
        lab, leaf = None, None
    assert hashlib.sha256(mask.tobytes()).hexdigest()[:16] == str(g["mask_digest"])
    assert hashlib.sha256(dep.tobytes()).hexdigest()[:16] == str(g["depth_digest"])
    return mask, dep, P, g, lab, leaf
