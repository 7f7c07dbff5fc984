"""OptimalLeafSelector drop-in (reference scripts/utils/leaf_scorer.py:10-207).

Same constructor, ``set_camera_params``, ``select_optimal_leaf(mask_tensor, depth_tensor) -> int | None``
and ``get_tall_leaves``; the work is one call into the native library (csrc/lg_stage1.cu).  Errors follow
the reference's convention: logged, ``None`` returned (leaf_scorer.py:201-203).
"""
from __future__ import annotations

import torch

from . import _log, _native as N
from .pipeline import GraspEngine, camera_from_projection


class OptimalLeafSelector:
    def __init__(self, device, max_labels: int = 128):
        N.lib()                      # fail loudly at construction if the CUDA library is missing
        self.device = device
        self.max_labels = max_labels
        self.camera_cx = None
        self.camera_cy = None
        self.f_norm = None
        self._engine = None
        self._tall_leaves = []
        self.last_records = None

    def set_camera_params(self, projection_matrix):
        self.f_norm = projection_matrix[0, 0]
        self.camera_cx = projection_matrix[0, 2]
        self.camera_cy = projection_matrix[1, 2]

    def _get_engine(self, h, w):
        e = self._engine
        if e is None or (e.H, e.W) != (h, w):
            dev = self.device if torch.device(self.device).type == "cuda" else None
            self._engine = e = GraspEngine(1, h, w, self.max_labels, device=dev)
        return e

    def select_optimal_leaf(self, mask_tensor, depth_tensor):
        try:
            if self.f_norm is None:
                raise ValueError("camera parameters not set")
            h, w = mask_tensor.shape[-2:]
            eng = self._get_engine(h, w)
            cam = N.Camera(float(self.f_norm), float(self.camera_cx), float(self.camera_cy))
            ids, rec = eng.select_leaf(mask_tensor, depth_tensor, cam)
            if ids[0] < 0:
                # "no leaf" is also what a label id outside the engine's tables produces (LG_ST_LABEL_RANGE); the
                # reference accepts any int16 id, so check (on this rare path only) and grow the tables
                mt = torch.as_tensor(mask_tensor)
                lo, top = (int(mt.min()), int(mt.max())) if mt.numel() else (0, 0)
                if lo < 0 or top >= 1024:
                    raise ValueError(f"label ids must lie in [0, 1024) (found {lo}..{top})")
                if top >= self.max_labels:
                    self.max_labels = min(1024, max(2 * self.max_labels, top + 1))
                    self._engine = None
                    eng = self._get_engine(h, w)
                    ids, rec = eng.select_leaf(mask_tensor, depth_tensor, cam)
            rec = rec[0]
            self.last_records = rec[rec["area"] > 0]
            self._tall_leaves = [int(r["leaf_id"]) for r in self.last_records if r["is_tall"]]
            _log.loginfo(f"Found {len(self._tall_leaves)} tall leaves")
            if ids[0] < 0:
                _log.logwarn("No valid leaf candidates found")
                return None
            return int(ids[0])
        except N.NativeError:
            raise
        except Exception as e:  # noqa: BLE001 - the reference swallows everything here
            _log.logerr(f"Error in leaf selection: {e}")
            return None

    def get_tall_leaves(self):
        return self._tall_leaves
