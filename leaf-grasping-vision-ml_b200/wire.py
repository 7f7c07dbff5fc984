"""Wire format either side of the path.

In:  the two ROS messages the node subscribes to (reference msg/masks.msg ``uint16[] imageData``, msg/depth.msg
     ``float32[] imageData``), flat row-major H x W, converted exactly as ``mask_callback`` / ``depth_callback`` do
     (scripts/leaf_grasp_node_v3.py:185-205: ``np.array(msg.imageData, dtype=np.int16)`` / ``dtype=np.float32``, then
     ``reshape(height, width)``).  ``stage_frames`` writes any number of such messages straight into one pinned
     host batch, which is what ``GraspEngine.process_batch_host`` (lg_process_batch_host) consumes.
Out: the comma-separated string ``publish_results`` sends on /optimal_leaf_grasp
     (scripts/leaf_grasp_node_v3.py:160-175).

No ROS dependency: a message is anything with an ``imageData`` sequence, or the sequence itself.
"""
from __future__ import annotations

import numpy as np
import torch


def _payload(msg):
    return getattr(msg, "imageData", msg)


def mask_from_wire(msg, height: int, width: int) -> np.ndarray:
    """uint16[] -> int16 [H, W] (ids above 32767 wrap negative exactly as in the reference's np.int16 conversion)."""
    a = np.asarray(_payload(msg))
    if a.dtype != np.int16:
        a = a.astype(np.uint16, copy=False).view(np.int16) if a.dtype == np.uint16 else a.astype(np.int64).astype(np.uint16).view(np.int16)
    if a.size != height * width:
        raise ValueError(f"mask message holds {a.size} values, expected {height}x{width}")
    return a.reshape(height, width)


def depth_from_wire(msg, height: int, width: int) -> np.ndarray:
    """float32[] -> float32 [H, W]."""
    a = np.asarray(_payload(msg), dtype=np.float32)
    if a.size != height * width:
        raise ValueError(f"depth message holds {a.size} values, expected {height}x{width}")
    return a.reshape(height, width)


def stage_frames(mask_msgs, depth_msgs, height: int, width: int, pin: bool = True):
    """Messages of n frames -> (labels int16 [n,H,W], depth float32 [n,H,W]) host tensors, pinned for the
    asynchronous copies of lg_process_batch_host."""
    n = len(mask_msgs)
    if len(depth_msgs) != n:
        raise ValueError("need one depth message per mask message")
    labels = torch.empty((n, height, width), dtype=torch.int16, pin_memory=pin and torch.cuda.is_available())
    depth = torch.empty((n, height, width), dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
    ln, dn = labels.numpy(), depth.numpy()
    for i in range(n):
        ln[i] = mask_from_wire(mask_msgs[i], height, width)
        dn[i] = depth_from_wire(depth_msgs[i], height, width)
    return labels, depth


def format_result(grasp_point_2d, grasp_point_3d, pre_grasp_point=None) -> str:
    """The String the node publishes (leaf_grasp_node_v3.py:168-173): 8 comma-separated values, 5 without a pre-grasp."""
    g2, g3 = grasp_point_2d, grasp_point_3d
    if pre_grasp_point is not None:
        p = pre_grasp_point
        return f"{g2[0]},{g2[1]},{g3[0]},{g3[1]},{g3[2]},{p[0]},{p[1]},{p[2]}"
    return f"{g2[0]},{g2[1]},{g3[0]},{g3[1]},{g3[2]}"


def format_frame_result(rec) -> str | None:
    """Same string from one lg_frame_result record (GraspEngine output); None when the frame has no grasp."""
    if int(rec["leaf_id"]) < 0 or int(rec["n_candidates"]) == 0:
        return None
    pre = rec["pre_grasp"]
    pre = None if np.isnan(pre).any() else tuple(float(v) for v in pre)
    return format_result((int(rec["grasp_x"]), int(rec["grasp_y"])), tuple(float(v) for v in rec["grasp_3d"]), pre)
