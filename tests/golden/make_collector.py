"""Golden vectors for the training-sample collector, produced by RUNNING THE REFERENCE's
EnhancedGraspDataCollector (scripts/utils/ml_grasp_optimizer/data_collector.py) in the build container:

    python tests/golden/make_collector.py

The reference draws from the unseeded global generators of `random` and `torch`.  To pin it, the module's
`random` and `torch.randn_like` are replaced by the counter-based generator of oracle.CollectorRng - the
reference code itself is unmodified and every other step (slicing, validation, rot90, noise arithmetic,
point rotation, the three negative-candidate sets, the attempt loop) is the reference's own.  The object is
built without __init__ so that nothing is written under ~/leaf_grasp_output.

Cases: the four SMALL synthetic frames (leaf and grasp point chosen by the reference), and three crafted
masks whose leaf reaches into the bottom quarter of the image and carries one-pixel spikes, so that the
stem and edge sets are not empty.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "leaf-grasping-vision-ml_b200"))
warnings.filterwarnings("ignore")

import ref_harness  # noqa: E402
import leafgrasp_oracle as O  # noqa: E402
import synth  # noqa: E402
from collector_cases import crafted, CONFIG_SEED, RNG_SEED  # noqa: E402



class ShimRandom:
    """Stands in for the `random` module inside data_collector.py."""

    def __init__(self, rng: O.CollectorRng, sets):
        self.rng, self.sets = rng, sets
        self.k = 0            # rotation being augmented (1..3)
        self.attempt = 0
        self.calls_in_attempt = 0

    def uniform(self, a, b):
        if (a, b) == (0.01, 0.02):
            self.k += 1
            return self.rng.noise_factor(self.k)
        assert (a, b) == (0.95, 1.0)
        return self.rng.score_jitter(self.k)

    def sample(self, pts, n):
        assert n == 1
        # the reference asks tip, stem, edge in this order and skips empty sets (data_collector.py:319-327)
        kinds = [i for i, s in enumerate(self.sets) if s]
        kind = kinds[self.calls_in_attempt]
        assert len(pts) == len(self.sets[kind])
        self.calls_in_attempt += 1
        out = [pts[self.rng.pick(self.attempt, kind, len(pts))]]
        if self.calls_in_attempt == len(kinds):
            self.attempt += 1
            self.calls_in_attempt = 0
        return out


def torch_proxy(shim: ShimRandom):
    proxy = types.ModuleType("torch_proxy")
    proxy.__dict__.update(torch.__dict__)
    proxy.randn_like = lambda t: torch.from_numpy(shim.rng.normal_patch(shim.k).copy())
    return proxy


def run_case(dc_mod, GPS, IP, mask_u8, depth, P, grasp, tag, frame_index):
    H, W = mask_u8.shape
    dev = torch.device("cpu")
    gps = GPS(dev)
    gps.set_camera_params(P)
    ip = IP(H, W, 21, 5)
    dt = torch.from_numpy(depth)
    scores = gps._calculate_all_scores(mask_u8, dt, ip)
    total = float(np.max(scores["traditional_score"]))
    if grasp is None:
        valid = gps._get_valid_regions(mask_u8, scores)
        cands = gps._get_candidate_points(scores["traditional_score"], valid, top_k=20, min_distance=10)
        grasp = tuple(int(v) for v in cands[0])
    col = object.__new__(dc_mod.EnhancedGraspDataCollector)
    col.patch_size = 32
    col.samples = []
    col.stats = {"positive_samples": 0, "negative_samples": 0, "augmented_samples": 0}
    col.save_samples = lambda: None
    col._log_collection_progress = lambda: None
    mask_bool = mask_u8.astype(bool)
    sets = (col._get_tip_points(mask_bool), col._get_stem_points(mask_bool), col._get_edge_points(mask_bool))
    rng = O.CollectorRng(RNG_SEED, frame_index)
    shim = ShimRandom(rng, sets)
    dc_mod.random = shim
    dc_mod.torch = torch_proxy(shim)
    ok = col.collect_sample(torch.from_numpy(mask_bool), dt, None, scores, grasp, total)
    dc_mod.random = __import__("random")
    dc_mod.torch = torch
    assert ok, tag
    # inputs are not stored: tests rebuild them from synth.make_frame / collector_cases.crafted
    out = {"grasp": np.array(grasp, np.int32), "total_score": np.float64(total),
           "frame_index": np.int32(frame_index), "rng_seed": np.int64(RNG_SEED),
           "mask_digest": np.array(hashlib.sha256(mask_u8.tobytes()).hexdigest()[:16]),
           "depth_digest": np.array(hashlib.sha256(depth.tobytes()).hexdigest()[:16])}
    for name, s in zip(("tip", "stem", "edge"), sets):
        out["set_" + name] = np.array([(int(x), int(y)) for (x, y) in s], np.int32).reshape(-1, 2)
    n = len(col.samples)
    out["patches"] = np.stack([torch.cat([s["depth_patch"].float()[None], s["mask_patch"].float()[None],
                                          torch.as_tensor(s["score_patches"]).float()]).numpy() for s in col.samples])
    out["labels"] = np.array([s["label"] for s in col.samples], np.int32)
    out["is_augmented"] = np.array([s["is_augmented"] for s in col.samples], np.int32)
    out["points"] = np.array([s["grasp_point"] for s in col.samples], np.int32)
    out["total_scores"] = np.array([s["total_score"] for s in col.samples], np.float64)
    # what the oracle says, checked here once against the reference before anything is stored
    mine = O.collect_sample(mask_u8, depth, scores, grasp, total, O.CollectorRng(RNG_SEED, frame_index))
    assert len(mine) == n, (tag, len(mine), n)
    for a, s in zip(mine, col.samples):
        assert a["label"] == s["label"] and a["is_augmented"] == s["is_augmented"], tag
        assert tuple(a["grasp_point"]) == tuple(s["grasp_point"]), (tag, a["grasp_point"], s["grasp_point"])
    print(tag, "samples", n, "sets", [len(s) for s in sets], "stats", col.stats)
    return out


def main():
    ref_harness.install()
    import scripts.utils.ml_grasp_optimizer.data_collector as dc_mod
    _, GPS, IP, _ = ref_harness.load()
    OLS = ref_harness.load()[0]
    cases = {}
    P = synth.projection_matrix(synth.SMALL)
    for idx in range(4):
        lab, dep = synth.make_frame(synth.SMALL, CONFIG_SEED, idx)
        ols = OLS(torch.device("cpu"))
        ols.set_camera_params(P)
        leaf = ols.select_optimal_leaf(torch.from_numpy(lab), torch.from_numpy(dep))
        mask = (lab == leaf).astype(np.uint8)
        cases[f"small{idx}"] = run_case(dc_mod, GPS, IP, mask, dep, P, None, f"small{idx}", idx)
        cases[f"small{idx}"]["leaf_id"] = np.int32(leaf)
    for idx in range(3):
        m, d = crafted(idx)
        cases[f"crafted{idx}"] = run_case(dc_mod, GPS, IP, m, d, P, None, f"crafted{idx}", 100 + idx)
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "collector.npz"), **flat)
    print("wrote collector.npz", os.path.getsize(os.path.join(HERE, "collector.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
