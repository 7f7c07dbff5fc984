"""B200-native grasp-selection hot path of Leaf-Grasping-Vision-ML.

Host side in Python mirroring the reference's entry points; every computation on the path runs in
the hand-written sm_100a kernels of ``csrc/`` behind the C-ABI declared in ``include/leafgrasp.h``.
There is no CPU fallback: using any compute entry point without the built library raises.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
