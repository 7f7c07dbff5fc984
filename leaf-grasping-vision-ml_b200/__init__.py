"""B200-native grasp-selection hot path of Leaf-Grasping-Vision-ML.

Host side in Python mirroring the reference's entry points; every computation on the path runs in
the hand-written sm_100a kernels of ``csrc/`` behind the C-ABI declared in ``include/leafgrasp.h``.
There is no CPU fallback: using any compute entry point without the built library raises.
"""
from . import synth, wire  # noqa: F401
from .data_collector import EnhancedGraspDataCollector  # noqa: F401
from .cnn import GraspPointCNN, fold_batchnorm, pack_weights  # noqa: F401
from .grasp_point_selector import GraspPointSelector  # noqa: F401
from .image_processor import ImageProcessor  # noqa: F401
from .leaf_scorer import OptimalLeafSelector  # noqa: F401
from .pipeline import GraspEngine, camera_from_projection  # noqa: F401

__all__ = ["synth", "wire", "GraspPointCNN", "GraspPointSelector", "ImageProcessor", "OptimalLeafSelector", "GraspEngine", "EnhancedGraspDataCollector",
           "camera_from_projection", "fold_batchnorm", "pack_weights"]
