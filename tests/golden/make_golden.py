"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run by hand in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

It imports the unmodified reference modules through ref_harness.py, feeds them the seeded synthetic
frames of <package>/synth.py and stores what they return.  The reference has no fixtures of its
own, so these files are the parity pin for oracle/leafgrasp_oracle.py.
Environment of the committed vectors: numpy 2.3.5, OpenCV 4.13.0 (IPP off), SciPy 1.18.1,
torch 2.11.0 CPU.  skfmm/paretoset are shimmed as described in ref_harness.py.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
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

MAP_KEYS = ("sdf_score", "approach_score", "flatness_map", "isolation_map", "distance_map",
            "accessibility_map", "stem_penalty", "traditional_score")
CONFIG_SEED = 7
CNN_SEED = 1234


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


class Recorder(torch.nn.Module):
    """Wraps the reference CNN and keeps every input it is called with."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.seen = []

    def forward(self, x):
        self.seen.append(x.detach().clone())
        return self.inner(x)


def run_frame(spec, idx, classes, sd, full_maps):
    OLS, GPS, IP, CNN = classes
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_frame(spec, CONFIG_SEED, idx)
    dev = torch.device("cpu")
    ols = OLS(dev)
    ols.set_camera_params(P)
    gps = GPS(dev)
    gps.set_camera_params(P)
    net = CNN(in_channels=9)
    net.load_state_dict(sd)
    net.eval()
    rec = Recorder(net)
    rec.eval()
    gps.ml_predictor = rec
    ip = IP(spec.height, spec.width, 21, 5)
    mt, dt = torch.from_numpy(lab), torch.from_numpy(dep)

    out = {"labels_digest": digest(lab), "depth_digest": digest(dep)}
    leaf = ols.select_optimal_leaf(mt, dt)
    out["leaf_id"] = -1 if leaf is None else int(leaf)
    out["tall"] = np.array(ols.get_tall_leaves(), dtype=np.int32)
    if leaf is None:
        return out
    om = mt == leaf
    mnp = om.numpy().astype(np.uint8)
    scores = gps._calculate_all_scores(mnp, dt, ip)
    valid = gps._get_valid_regions(mnp, scores)
    cands = gps._get_candidate_points(scores["traditional_score"], valid, top_k=20, min_distance=10)
    key = scores["traditional_score"] * valid
    out["candidates"] = np.array(cands, dtype=np.int32)
    out["n_positive"] = int(sum(1 for (x, y) in cands if key[y, x] > 0))
    out["trad_at"] = np.array([scores["traditional_score"][y, x] for (x, y) in cands], dtype=np.float64)
    out["valid_digest"] = digest(valid)
    out["valid_count"] = int(valid.sum())
    ang, major, minor, center = gps.estimate_leaf_orientation(mnp)
    out["angle"] = np.float64(ang)
    rng = np.random.default_rng(1000 + idx)
    ys, xs = np.nonzero(mnp)
    pick = rng.choice(len(ys), size=min(256, len(ys)), replace=False)
    out["sample_yx"] = np.stack([ys[pick], xs[pick]], axis=1).astype(np.int32)
    for k in MAP_KEYS:
        m = np.asarray(scores[k])
        out["digest_" + k] = digest(m)
        out["sample_" + k] = m[ys[pick], xs[pick]].astype(np.float64)
    if full_maps:
        y0, y1 = max(0, ys.min() - 8), min(spec.height, ys.max() + 9)
        x0, x1 = max(0, xs.min() - 8), min(spec.width, xs.max() + 9)
        out["crop"] = np.array([y0, y1, x0, x1], dtype=np.int32)
        for k in MAP_KEYS:
            out["map_" + k] = np.asarray(scores[k])[y0:y1, x0:x1].astype(np.float32)
        out["map_valid"] = valid[y0:y1, x0:x1]
        out["map_dist_q16"] = O.chamfer5_q16(mnp)[y0:y1, x0:x1]
    rec.seen.clear()
    g2, g3, gpre = gps.select_grasp_point(om, dt, ip)
    out["grasp_2d"] = np.array(g2, dtype=np.int32)
    out["grasp_3d"] = np.array(g3, dtype=np.float64)
    out["pre_grasp"] = np.array(gpre, dtype=np.float64)
    if rec.seen:
        feats = torch.cat(rec.seen, dim=0)
        with torch.no_grad():
            logits = net(feats).reshape(-1)
        out["n_ml"] = feats.shape[0]
        out["logits"] = logits.numpy().astype(np.float32)
        out["patch_digest"] = digest(feats.numpy())
        if full_maps:
            out["patches"] = feats.numpy().astype(np.float32)
    else:
        out["n_ml"] = 0
    return out


def main():
    classes = ref_harness.load()
    sd = O.seeded_state_dict(CNN_SEED)
    meta = {"config_seed": CONFIG_SEED, "cnn_seed": CNN_SEED, "frames": []}
    plan = [("SMALL", 0, True), ("SMALL", 1, False), ("SMALL", 2, False), ("SMALL", 3, False),
            ("CFG1", 0, False), ("CFG2", 0, False), ("CFG2", 1, False)]
    for name, idx, full in plan:
        spec = getattr(synth, name)
        out = run_frame(spec, idx, classes, sd, full)
        fn = f"frame_{name.lower()}_{idx}.npz"
        np.savez_compressed(os.path.join(HERE, fn), **out)
        meta["frames"].append({"spec": name, "index": idx, "file": fn, "leaf_id": out["leaf_id"]})
        print(fn, "leaf", out["leaf_id"], "n_pos", out.get("n_positive"), "n_ml", out.get("n_ml"))

    # CNN-only vectors: the reference module on seeded random patches (BASELINE config 4 shape)
    _, _, _, CNN = classes
    net = CNN(in_channels=9)
    net.load_state_dict(sd)
    net.eval()
    g = torch.Generator().manual_seed(99)
    x = torch.rand(16, 9, 32, 32, generator=g)
    x[:, 1] = (x[:, 1] > 0.5).float()
    with torch.no_grad():
        y = net(x).reshape(-1)
        yo = O.cnn_forward(sd, x).reshape(-1)
    assert torch.allclose(y, yo, atol=1e-5), (y - yo).abs().max()
    np.savez_compressed(os.path.join(HERE, "cnn_patches.npz"), x=x.numpy(), logits=y.numpy())
    with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


if __name__ == "__main__":
    main()
