"""GraspPointSelector drop-in (reference scripts/utils/grasp_point_selector.py:13-826, live methods only).

Method names, argument order and return conventions follow the reference; each method is one call into
the native library.  ``image_processor`` is accepted for signature compatibility (the Gaussian / Sobel
stencils are fused into the score kernel).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _log, _native as N
from .cnn import GraspPointCNN
from .pipeline import GraspEngine

SCORE_KEYS = ("sdf_score", "approach_score", "flatness_map", "isolation_map", "distance_map", "accessibility_map",
              "stem_penalty")


class GraspPointSelector:
    def __init__(self, device):
        N.lib()
        self.device = device
        self.flatness_weight, self.isolation_weight = 0.25, 0.4         # unused by the reference too (:18-21)
        self.edge_weight, self.accessibility_weight = 0.2, 0.15
        self.min_flat_area_size, self.min_edge_distance, self.isolation_radius = 15, 20, 50
        self.camera_cx, self.camera_cy, self.f_norm = 707, 494, None      # :29-31
        self.baseline = None
        self.use_bf16_cnn = False
        self._engine = None
        self._engine_model_version = None
        self.last_result = None
        self.ml_predictor = GraspPointCNN(in_channels=9)
        self.load_ml_model()

    # ---- model -------------------------------------------------------------------------------------------
    def load_ml_model(self):
        """~/leaf_grasp_output/ml_models/best_model.pth, else CV-only (reference :43-57)."""
        try:
            path = os.path.expanduser("~/leaf_grasp_output/ml_models/best_model.pth")
            if os.path.exists(path):
                ckpt = torch.load(path, map_location="cpu")
                self.ml_predictor.load_state_dict(ckpt["model_state_dict"])
                self.ml_predictor.eval()
                _log.loginfo("Loaded ML grasp model")
            else:
                _log.logwarn("No ML model found, will use traditional scoring only")
                self.ml_predictor = None
        except Exception as e:  # noqa: BLE001
            _log.logerr(f"Error loading ML model: {e}")
            self.ml_predictor = None

    def set_camera_params(self, projection_matrix):
        self.f_norm = projection_matrix[0, 0]
        self.camera_cx = projection_matrix[0, 2]
        self.camera_cy = projection_matrix[1, 2]
        self.baseline = -projection_matrix[0, 3] / self.f_norm

    # ---- engine ------------------------------------------------------------------------------------------
    def _cam(self):
        if self.f_norm is None:
            raise ValueError("camera parameters not set (f_norm is None)")
        return N.Camera(float(self.f_norm), float(self.camera_cx), float(self.camera_cy))

    def _bf16(self) -> bool:
        return bool(self.use_bf16_cnn and self.ml_predictor is not None and self.ml_predictor.is_default_architecture)

    def _get_engine(self, h, w):
        e = self._engine
        if e is None or (e.H, e.W) != (h, w):
            dev = self.device if torch.device(self.device).type == "cuda" else None
            self._engine = e = GraspEngine(1, h, w, 2, device=dev)
            self._engine_model_version = None
        ver = None if self.ml_predictor is None else self.ml_predictor._version()
        if ver != self._engine_model_version:
            m = self.ml_predictor
            if m is None:
                e.set_cnn_weights(None)
            else:       # any architecture the reference's constructor accepts; only the default one has a bf16 path
                e.set_cnn_weights(m.packed(), None if m.is_default_architecture else m._config)
            self._engine_model_version = ver
        return e

    # ---- the reference's entry points --------------------------------------------------------------------
    def select_grasp_point(self, leaf_mask, depth_tensor, image_processor=None, pcl_data=None):
        try:
            h, w = leaf_mask.shape[-2:]
            eng = self._get_engine(h, w)
            r = eng.select_grasp_point(torch.as_tensor(leaf_mask).to(torch.uint8), torch.as_tensor(depth_tensor),
                                       self._cam(), self._bf16())[0]
            self.last_result = r
            if r["n_candidates"] == 0:
                _log.logwarn("No valid candidate points found")
                return None, None, None
            g2 = (int(r["grasp_x"]), int(r["grasp_y"]))
            g3 = tuple(float(v) for v in r["grasp_3d"])
            pre = tuple(float(v) for v in r["pre_grasp"])
            if any(np.isnan(v) for v in pre):     # depth 0 / NaN at the pick: the reference's projection raises inside
                pre = None                        # calculate_pre_grasp_point, which then returns None (:754-819)
            return g2, g3, pre
        except N.NativeError:
            raise
        except Exception as e:  # noqa: BLE001
            _log.logerr(f"Error in grasp point selection: {e}")
            return None, None, None

    def _calculate_all_scores(self, leaf_mask_np, depth_tensor, image_processor=None):
        h, w = leaf_mask_np.shape
        eng = self._get_engine(h, w)
        m = eng.score_maps(torch.as_tensor(np.ascontiguousarray(leaf_mask_np)).to(torch.uint8),
                           torch.as_tensor(depth_tensor), self._cam())
        out = {k: m[k][0].cpu().numpy() for k in SCORE_KEYS + ("traditional_score",)}
        out["_valid"] = m["valid"][0].cpu().numpy().astype(bool)
        return out

    def _get_valid_regions(self, leaf_mask_np, scores):
        if "_valid" in scores:
            return scores["_valid"]
        dev = torch.device("cuda")
        v = ((torch.as_tensor(scores["distance_map"], device=dev) > self.min_edge_distance)
             & (torch.as_tensor(leaf_mask_np, device=dev) > 0) & (torch.as_tensor(scores["stem_penalty"], device=dev) < 0.8))
        return v.cpu().numpy()

    def _get_candidate_points(self, score_map, valid_regions, top_k=20, min_distance=10):
        try:
            if top_k != 20 or min_distance != 10:
                raise NotImplementedError("the CUDA candidate search is built for top_k=20, min_distance=10 (:197-198)")
            h, w = score_map.shape
            eng = self._get_engine(h, w)
            xy, cnt = eng.candidate_points(torch.as_tensor(np.ascontiguousarray(score_map, dtype=np.float64)),
                                           torch.as_tensor(np.ascontiguousarray(valid_regions)).to(torch.uint8))
            n = int(cnt[0])
            return [(int(x), int(y)) for x, y in xy[0, :n].cpu().numpy()]
        except N.NativeError:
            raise
        except Exception as e:  # noqa: BLE001
            _log.logerr(f"Error getting candidate points: {e}")
            return []

    def get_ml_score(self, leaf_mask, depth_tensor, scores, point):
        """Window extraction is indexing on the host; normalisation and the CNN run on the device."""
        try:
            if self.ml_predictor is None:
                return None
            x, y = point
            h, w = depth_tensor.shape[-2:]
            ys = np.clip(np.arange(y - 16, y + 16), 0, h - 1)
            xs = np.clip(np.arange(x - 16, x + 16), 0, w - 1)
            mask_t = torch.as_tensor(leaf_mask)
            if mask_t.dtype == torch.bool and (x - 16 < 0 or y - 16 < 0 or x + 16 > w or y + 16 > h):
                return None          # the reference's replicate pad rejects bool tensors (:424-428)
            take = lambda a: torch.as_tensor(np.asarray(torch.as_tensor(a).cpu())[np.ix_(ys, xs)]).float()
            ch = [take(depth_tensor), take(mask_t)] + [take(scores[k]) for k in SCORE_KEYS]
            eng = self._get_engine(h, w)
            feats = eng.normalize_patches(torch.stack(ch)[None])
            logit = eng.cnn_forward(feats, self._bf16())
            s = torch.sigmoid(logit).item()
            return float(np.tanh(s * 3.0) * 0.5 + 0.5)
        except N.NativeError:
            raise
        except Exception as e:  # noqa: BLE001
            _log.logerr(f"Error in ML scoring: {e}")
            return None

    def estimate_leaf_orientation(self, leaf_mask_np):
        try:
            h, w = leaf_mask_np.shape
            o = self._get_engine(h, w).leaf_orientation(
                torch.as_tensor(np.ascontiguousarray(leaf_mask_np)).to(torch.uint8))[0].cpu().numpy()
            if np.isnan(o[0]):
                return None, None, None, None
            return float(o[0]), float(o[1]), float(o[2]), (float(o[3]), float(o[4]))
        except N.NativeError:
            raise
        except Exception as e:  # noqa: BLE001
            _log.logerr(f"Error in leaf orientation estimation: {e}")
            return None, None, None, None

    def get_3d_grasp_point(self, grasp_point_2d, depth_tensor, pcl_data=None):
        u, v = grasp_point_2d
        z = depth_tensor[v, u].item()
        return ((z * (u - self.camera_cx)) / self.f_norm, (z * (v - self.camera_cy)) / self.f_norm, z)
