"""Import the UNMODIFIED reference modules of the hot path in the build container.

Only ``make_golden.py`` (run by hand here, never on the GPU box) uses this: /root/reference does not
exist anywhere else.  What it installs before importing:
  * a ``rospy`` stub with the four logging calls the modules use;
  * ``skfmm.distance`` -> exact Euclidean distance transform (scikit-fmm 2022.3.26 is not installable
    here; that sub-result is parity-unpinned, see oracle/leafgrasp_oracle.py);
  * ``paretoset.paretoset`` -> non-dominated rows, first duplicate kept (paretoset 1.2.3 semantics);
  * cv2 IPP off, so distanceTransform is OpenCV's integer chamfer.
"""
from __future__ import annotations

import sys
import types

import cv2
import numpy as np
import scipy.ndimage as ndi

REFERENCE_ROOT = "/root/reference"


def install():
    cv2.ipp.setUseIPP(False)
    rospy = types.ModuleType("rospy")
    for name in ("loginfo", "logwarn", "logerr", "logdebug"):
        setattr(rospy, name, lambda *a, **k: None)
    sys.modules["rospy"] = rospy

    skfmm = types.ModuleType("skfmm")

    def distance(phi, dx=1):
        return ndi.distance_transform_edt(np.asarray(phi) != 0) * dx

    skfmm.distance = distance
    sys.modules["skfmm"] = skfmm

    pareto = types.ModuleType("paretoset")

    def paretoset(costs, sense=None, distinct=True):
        c = np.asarray(costs, dtype=np.float64)
        n = c.shape[0]
        keep = np.ones(n, dtype=bool)
        for i in range(n):
            for j in range(n):
                if i == j:
                    continue
                ge = np.all(c[j] >= c[i])
                gt = np.any(c[j] > c[i])
                if ge and (gt or (distinct and j < i)):
                    keep[i] = False
                    break
        return keep

    pareto.paretoset = paretoset
    sys.modules["paretoset"] = pareto
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load():
    """Returns (OptimalLeafSelector, GraspPointSelector, ImageProcessor, GraspPointCNN) classes."""
    install()
    from scripts.utils.leaf_scorer import OptimalLeafSelector
    from scripts.utils.grasp_point_selector import GraspPointSelector
    from scripts.utils.image_processor import ImageProcessor
    from scripts.utils.ml_grasp_optimizer.model import GraspPointCNN
    return OptimalLeafSelector, GraspPointSelector, ImageProcessor, GraspPointCNN
