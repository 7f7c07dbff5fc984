"""ctypes binding of liblgb200.so (include/leafgrasp.h).

The library is built in-tree by ``__graft_entry__.build()`` (``make -C csrc``).  There is no fallback:
if it is missing, ``lib()`` raises, and so does every compute call made without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblgb200.so")
TOP_K, PATCH, CHANNELS = 20, 32, 9

ST_NO_LEAF, ST_LABEL_RANGE, ST_RUNS_OVERFLOW, ST_NO_CANDIDATE = 1, 2, 4, 8


class CnnConfig(C.Structure):
    """lg_cnn_config: n_blocks, filters[4], attention (0 none, 1 spatial, 2 channel, 3 hybrid)."""
    _fields_ = [("n_blocks", C.c_int32), ("filters", C.c_int32 * 4), ("attention", C.c_int32)]


ATTENTION_CODES = {"none": 0, "spatial": 1, "channel": 2, "hybrid": 3}


def cnn_config(attention_type="spatial", encoder_filters=(64, 128, 256)) -> CnnConfig:
    f = list(encoder_filters)
    if attention_type not in ATTENTION_CODES or not 1 <= len(f) <= 4:
        raise ValueError(f"unsupported GraspPointCNN architecture: {attention_type}, {f}")
    return CnnConfig(len(f), (C.c_int32 * 4)(*(f + [0] * (4 - len(f)))), ATTENTION_CODES[attention_type])


class Camera(C.Structure):
    _fields_ = [("f", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


LEAF_RECORD = np.dtype([
    ("leaf_id", np.int32), ("area", np.uint32), ("median_depth", np.float32), ("mean_depth", np.float32),
    ("centroid_x", np.float64), ("centroid_y", np.float64), ("clutter", np.float64), ("distance", np.float64),
    ("visibility", np.float64), ("mean_distance", np.float64), ("is_tall", np.int32), ("is_candidate", np.int32),
], align=True)

FRAME_RESULT = np.dtype([
    ("status", np.uint32), ("leaf_id", np.int32), ("n_candidates", np.int32), ("n_positive", np.int32),
    ("cand_x", np.int32, (TOP_K,)), ("cand_y", np.int32, (TOP_K,)), ("trad", np.float64, (TOP_K,)),
    ("logit", np.float32, (TOP_K,)), ("ml", np.float64, (TOP_K,)), ("ml_valid", np.int32, (TOP_K,)),
    ("best_index", np.int32), ("ml_used", np.int32), ("best_score", np.float64),
    ("grasp_x", np.int32), ("grasp_y", np.int32), ("grasp_3d", np.float64, (3,)), ("pre_grasp", np.float64, (3,)),
    ("angle", np.float64), ("sdf_max", np.float32), ("region", np.int32, (4,)),
], align=True)

SAMPLES_PER_FRAME = 7
SAMPLE_KINDS = ("positive", "rot90", "rot180", "rot270", "tip", "stem", "edge")
SAMPLE_META = np.dtype([
    ("valid", np.int32), ("label", np.int32), ("is_augmented", np.int32), ("kind", np.int32),
    ("x", np.int32), ("y", np.int32), ("total_score", np.float64),
], align=True)

# every symbol include/leafgrasp.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "lg_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int]),
    "lg_destroy": (None, [_P]),
    "lg_last_error": (C.c_char_p, []),
    "lg_context_bytes": (C.c_uint64, [_P]),
    "lg_set_cnn_weights": (C.c_int, [_P, _P, C.c_uint64]),
    "lg_set_cnn_model": (C.c_int, [_P, C.POINTER(CnnConfig), _P, C.c_uint64]),
    "lg_cnn_model_floats": (C.c_uint64, [C.POINTER(CnnConfig)]),
    "lg_process_batch": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Camera), _P, C.c_int, _P]),
    "lg_process_batch_host": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Camera), _P, C.c_int, _P]),
    "lg_select_leaf": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Camera), _P, _P, _P]),
    "lg_chamfer_transform": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "lg_edt_squared": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "lg_score_maps": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Camera)] + [_P] * 10 + [_P]),
    "lg_candidate_points": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P]),
    "lg_cnn_forward": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P]),
    "lg_cnn_bf16_features": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "lg_select_grasp_point": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(Camera), _P, C.c_int, _P]),
    "lg_leaf_orientation": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "lg_collect_samples": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "lg_collector_points": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "lg_sizeof_sample_meta": (C.c_uint64, []),
    "lg_patches": (C.c_int, [_P, _P, C.c_int, _P]),
    "lg_normalize_patches": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "lg_smooth_depth": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "lg_set_record_output": (C.c_int, [_P, _P]),
    "lg_set_profiling": (C.c_int, [_P, C.c_int]),
    "lg_stage_times": (C.c_int, [_P, _P, C.c_int]),
    "lg_stage_times_mean": (C.c_int, [_P, _P, C.c_int, _P]),
    "lg_set_overlap": (C.c_int, [_P, C.c_int]),
    "lg_set_patch_export": (C.c_int, [_P, C.c_int]),
    "lg_host_memory_is_pinned": (C.c_int, [_P]),
    "lg_set_host_label_rle": (C.c_int, [_P, C.c_int]),
    "lg_host_call_bytes": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "lg_rle_encode_labels": (C.c_uint32, [_P, C.c_int, C.c_int, _P, C.c_uint32, _P, C.c_int]),
    "lg_rle_host_isa": (C.c_int, []),
    "lg_set_score_weights": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double]),
    "lg_launch_count": (C.c_uint64, []),
    "lg_sizeof_frame_result": (C.c_uint64, []),
    "lg_sizeof_leaf_record": (C.c_uint64, []),
    "lg_cnn_weight_floats": (C.c_uint64, []),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load liblgb200.so once; raise if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C leaf-grasping-vision-ml_b200/csrc).  This package has no CPU fallback.")
    h = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(h, name)      # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if (h.lg_sizeof_frame_result() != FRAME_RESULT.itemsize or h.lg_sizeof_leaf_record() != LEAF_RECORD.itemsize or
            h.lg_sizeof_sample_meta() != SAMPLE_META.itemsize):
        raise NativeError("struct layout mismatch between _native.py and include/leafgrasp.h")
    _lib = h
    return h


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().lg_last_error()
        raise NativeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
