"""TEST INFRASTRUCTURE - not part of the product.  ctypes front end of oracle/fmm_distance.c, the restatement of
scikit-fmm 2022.3.26 `skfmm.distance` (reference scripts/utils/leaf_scorer.py:69).  Parity unpinned: scikit-fmm is not
installed here, the restatement follows its published algorithm (see the C file's header)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfmm_oracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "fmm_distance.c")):
            subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = C.CDLL(_SO)
        _lib.fmm_distance_2d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.fmm_distance_2d.restype = C.c_int
    return _lib


def distance(phi: np.ndarray) -> np.ndarray:
    """skfmm.distance(phi, dx=1) for a 2-D array; raises ValueError when phi has no zero level set, as skfmm does."""
    p = np.ascontiguousarray(phi, dtype=np.float64)
    out = np.empty_like(p)
    rc = _load().fmm_distance_2d(p.ctypes.data, p.shape[0], p.shape[1], out.ctypes.data)
    if rc == -1:
        raise ValueError("the array phi contains no zero contour (no zero level set)")
    if rc != 0:
        raise MemoryError("fmm_distance_2d")
    return out


def negative_discriminants() -> int:
    """How often the quadratic of the update had no real root since the library was loaded."""
    return int(C.c_longlong.in_dll(_load(), "fmm_negative_discriminants").value)


def clutter_extrema_fmm(labels: np.ndarray):
    """leaf_scorer.py:67-71 with the fast-marching field: ((row, col) of argmin, (row, col) of argmax)."""
    phi = np.where(np.asarray(labels) >= 1, 0, 1)
    field = distance(phi)
    pmin = np.unravel_index(field.argmin(), field.shape)
    pmax = np.unravel_index(field.argmax(), field.shape)
    return (int(pmin[0]), int(pmin[1])), (int(pmax[0]), int(pmax[1]))
