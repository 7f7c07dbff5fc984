"""Training-sample collector: host-side mirror of the reference's EnhancedGraspDataCollector
(scripts/utils/ml_grasp_optimizer/data_collector.py) over the device kernel of csrc/lg_collect.cu.

Same sample dictionaries, statistics, save cadence and `training_data.pt` layout as the reference, so
`ml_grasp_optimizer/dataset.py` and `train_model.py` read what this class writes.  What differs:

* the patches come from `GraspEngine.collect_samples` (all frames of a batch in one launch) instead of Python slicing;
* random draws (negative picks, depth noise, score jitter) are a pure function of (seed, frame counter) instead of the
  process-global generators - the same samples come out on every run;
* `data_dir` is an argument (default: the reference's ~/leaf_grasp_output/ml_training_data).
"""
from __future__ import annotations

import os
import shutil

import numpy as np
import torch

from . import _native as N

SCORE_NAMES = ("sdf_score", "approach_score", "flatness_map", "isolation_map", "distance_map", "accessibility_map",
               "stem_penalty")                                     # data_collector.py:131-133


class EnhancedGraspDataCollector:
    def __init__(self, patch_size: int = 32, resume: bool = True, data_dir: str | None = None, engine=None, seed: int = 0):
        if patch_size != N.PATCH:
            raise ValueError("the device path extracts 32x32 windows (the size every reference caller uses)")
        self.patch_size = patch_size
        self.samples = []
        self.data_dir = data_dir or os.path.expanduser("~/leaf_grasp_output/ml_training_data")
        if not resume and os.path.exists(self.data_dir):           # :19-21
            shutil.rmtree(self.data_dir)
        os.makedirs(self.data_dir, exist_ok=True)
        self.stats = {"positive_samples": 0, "negative_samples": 0, "augmented_samples": 0}
        self.engine = engine
        self.seed = int(seed)
        self.frames_seen = 0          # frame counter of the draw function
        self._cam = None
        if resume:
            self.load_existing_data()

    # ---- reference-shaped entry point -------------------------------------------------------------------
    def set_camera_params(self, P):
        from .pipeline import camera_from_projection
        self._cam = camera_from_projection(P)

    def collect_sample(self, leaf_mask, depth_tensor, rgb_image, scores, grasp_point_2d, total_score) -> bool:
        """data_collector.py:175.  `scores` is accepted for signature compatibility only: the device recomputes the maps
        from (leaf_mask, depth_tensor), which is what the reference's caller computed them from."""
        if self.engine is None or self._cam is None:
            raise N.NativeError("collect_sample needs an engine and set_camera_params(P)")
        if not torch.is_tensor(depth_tensor) or not torch.is_tensor(leaf_mask):
            return False
        mask = leaf_mask.to(torch.uint8)
        self.engine.select_grasp_point(mask, depth_tensor, self._cam)
        patches, meta, _ = self.engine.collect_samples(depth_tensor, mask=mask, seed=self.seed,
                                                       first_frame_index=self.frames_seen,
                                                       grasp_xy=[[int(grasp_point_2d[0]), int(grasp_point_2d[1])]],
                                                       total_score=[float(total_score)])
        return self.collect_batch(patches, meta) == 1

    # ---- batched entry point ------------------------------------------------------------------------------
    def collect_from_engine(self, depth, labels=None, mask=None) -> int:
        """After engine.process_batch(labels, depth, ...) or engine.select_grasp_point(mask, depth, ...): one
        collect_sample per frame with the grasp point just selected.  Returns the number of frames that gave samples."""
        patches, meta, _ = self.engine.collect_samples(depth, labels=labels, mask=mask, seed=self.seed,
                                                       first_frame_index=self.frames_seen)
        return self.collect_batch(patches, meta)

    def collect_batch(self, patches, meta) -> int:
        """patches [n,7,9,32,32] (tensor or ndarray), meta [n,7] SAMPLE_META: append the valid slots frame by frame in the
        reference's order (positive, three rotations, negatives) and keep its save cadence (:236-238)."""
        p = torch.as_tensor(patches).detach().cpu()
        n = p.shape[0]
        ok_frames = 0
        for b in range(n):
            self.frames_seen += 1
            if not meta[b, 0]["valid"]:
                continue                                           # collect_sample returned False: nothing was added
            ok_frames += 1
            for k in range(N.SAMPLES_PER_FRAME):
                m = meta[b, k]
                if not m["valid"]:
                    continue
                self._add_sample(p[b, k, 0].clone(), p[b, k, 1].clone(), p[b, k, 2:].clone(), float(m["total_score"]),
                                 (int(m["x"]), int(m["y"])), int(m["label"]), bool(m["is_augmented"]))
            if (self.stats["positive_samples"] + self.stats["negative_samples"]) % 5 == 0:
                self.save_samples()
        return ok_frames

    def _add_sample(self, depth_patch, mask_patch, score_patches, total_score, grasp_point, label, is_augmented) -> bool:
        """data_collector.py:350-395."""
        size = (self.patch_size, self.patch_size)
        if tuple(depth_patch.shape) != size or tuple(mask_patch.shape) != size:
            return False
        self.samples.append({
            "depth_patch": depth_patch, "mask_patch": mask_patch.float(), "score_patches": score_patches,
            "total_score": float(total_score), "grasp_point": tuple(map(int, grasp_point)), "label": int(label),
            "is_augmented": bool(is_augmented)})
        if label == 1:
            self.stats["augmented_samples" if is_augmented else "positive_samples"] += 1
        else:
            self.stats["negative_samples"] += 1
        return True

    # ---- persistence (data_collector.py:42-81, 497-598) -----------------------------------------------------
    def _tensors(self):
        s = self.samples
        return {
            "depth_patches": torch.stack([x["depth_patch"] for x in s]),
            "mask_patches": torch.stack([x["mask_patch"] for x in s]),
            "score_patches": torch.stack([x["score_patches"] for x in s]),
            "labels": torch.tensor([x["label"] for x in s]),
            "total_scores": torch.tensor([x["total_score"] for x in s]),
            "grasp_points": torch.tensor([x["grasp_point"] for x in s]),
            "is_augmented": torch.tensor([x["is_augmented"] for x in s]),
        }

    def quality_metrics(self, data=None) -> dict:
        d = data or self._tensors()
        ts = d["total_scores"]
        return {
            "depth_range": [d["depth_patches"].min().item(), d["depth_patches"].max().item()],
            "mask_coverage": (d["mask_patches"] > 0).float().mean().item(),
            "positive_ratio": (d["labels"] == 1).float().mean().item(),
            "augmented_ratio": d["is_augmented"].float().mean().item(),
            "score_statistics": {"mean": ts.mean().item(), "std": ts.std().item(), "min": ts.min().item(),
                                 "max": ts.max().item()},
        }

    def save_samples(self):
        if not self.samples:
            return
        save_path = os.path.join(self.data_dir, "training_data.pt")
        backup = save_path + ".backup"
        if os.path.exists(save_path):
            shutil.copy2(save_path, backup)
        try:
            data = self._tensors()
            q = self.quality_metrics(data)
            torch.save(data, save_path)
            with open(os.path.join(self.data_dir, "collection_metadata.txt"), "w") as f:
                f.write("=== Data Collection Statistics ===\n")
                f.write(f"Original positive samples: {self.stats['positive_samples']}\n")
                f.write(f"Augmented positive samples: {self.stats['augmented_samples']}\n")
                f.write(f"Negative samples: {self.stats['negative_samples']}\n")
                f.write(f"Total samples: {len(self.samples)}\n\n=== Tensor Shapes ===\n")
                for key, t in data.items():
                    f.write(f"{key}: {t.shape}\n")
                f.write("\n=== Quality Metrics ===\n")
                f.write(f"Depth range: {q['depth_range']}\n")
                f.write(f"Mask coverage: {q['mask_coverage']:.3f}\n")
                f.write(f"Positive ratio: {q['positive_ratio']:.3f}\n")
                f.write(f"Augmented ratio: {q['augmented_ratio']:.3f}\n\nScore Statistics:\n")
                for key, v in q["score_statistics"].items():
                    f.write(f"{key}: {v:.3f}\n")
            if os.path.exists(backup):
                os.remove(backup)
        except Exception:
            if os.path.exists(backup):
                shutil.copy2(backup, save_path)
            raise
        with open(os.path.join(self.data_dir, "collection_progress.txt"), "w") as f:
            f.write(f"last_frame: {self.stats['positive_samples']}\n")

    def load_existing_data(self):
        save_path = os.path.join(self.data_dir, "training_data.pt")
        if not os.path.exists(save_path):
            return
        data = torch.load(save_path)
        for i in range(len(data["labels"])):
            self.samples.append({
                "depth_patch": data["depth_patches"][i], "mask_patch": data["mask_patches"][i],
                "score_patches": data["score_patches"][i], "total_score": data["total_scores"][i].item(),
                "grasp_point": tuple(data["grasp_points"][i].tolist()), "label": data["labels"][i].item(),
                "is_augmented": data["is_augmented"][i].item()})
        s = self.samples
        self.stats["positive_samples"] = sum(1 for x in s if x["label"] == 1 and not x["is_augmented"])
        self.stats["augmented_samples"] = sum(1 for x in s if x["label"] == 1 and x["is_augmented"])
        self.stats["negative_samples"] = sum(1 for x in s if x["label"] == 0)
        self.frames_seen = self.stats["positive_samples"]
