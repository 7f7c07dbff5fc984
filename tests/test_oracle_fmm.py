"""The restatement of scikit-fmm's 2nd-order fast-marching distance (oracle/fmm_distance.c): the field the reference
computes at leaf_scorer.py:69 and of which it uses only the arg-max (:71).  scikit-fmm is not installed, so these tests pin
the restatement's own properties and its relation to the exact Euclidean transform the rest of the repository uses; the
arg-max agreement over >= 100 frames per configuration is in profiles/r4/fmm_vs_edt.json (tools/fmm_vs_edt.py)."""
import os
import sys

import numpy as np
import pytest
import scipy.ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import fmm  # noqa: E402
import leafgrasp_oracle as O  # noqa: E402
from leafgrasp_b200 import synth  # noqa: E402


def test_no_zero_level_set_raises_like_skfmm():
    with pytest.raises(ValueError):
        fmm.distance(np.ones((8, 9)))


def test_straight_edge_second_order_values():
    """Sources filling the left half plane: the field depends on the column only.  From cell-centred zeros the 2nd-order
    one-sided difference (3 t - 4 v1 + v2) / 2 = 1 gives 2/3 in the first background column and then t = tp + 2/3 with
    tp = (4 v1 - v2) / 3: 14/9, 68/27, ... - the hand-computed recurrence of the published update rule."""
    phi = np.ones((12, 20))
    phi[:, :6] = 0
    d = fmm.distance(phi)
    assert np.all(d[:, :6] == 0)
    want = [2.0 / 3.0]
    v2, v1 = 0.0, want[0]
    for _ in range(5):
        t = (4 * v1 - v2) / 3 + 2.0 / 3.0
        want.append(t)
        v2, v1 = v1, t
    for k, w in enumerate(want):
        np.testing.assert_allclose(d[:, 6 + k], w, rtol=0, atol=1e-12)


def test_field_close_to_exact_transform_and_monotone_from_sources():
    lab, _ = synth.make_frame(synth.SMALL, 7, 1)
    leafy = lab >= 1
    f = fmm.distance(np.where(leafy, 0, 1))
    e = ndi.distance_transform_edt(~leafy)
    assert np.all(f[leafy] == 0) and np.all(f[~leafy] > 0)
    assert np.abs(f - e).max() < 1.0          # within a pixel of the exact distance everywhere
    # every background pixel has a 4-neighbour that is closer to the leaves (the marcher's causality)
    pad = np.pad(f, 1, mode="constant", constant_values=np.inf)
    nmin = np.minimum(np.minimum(pad[:-2, 1:-1], pad[2:, 1:-1]), np.minimum(pad[1:-1, :-2], pad[1:-1, 2:]))
    assert np.all(nmin[~leafy] < f[~leafy])


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_argmax_and_selected_leaf_agree_with_exact_transform_on_small_frames(idx):
    spec = synth.SMALL
    P = synth.projection_matrix(spec)
    lab, dep = synth.make_frame(spec, 7, idx)
    pmin_f, pmax_f = fmm.clutter_extrema_fmm(lab)
    pmin_e, pmax_e = O.clutter_extrema(lab)
    assert pmin_f == pmin_e                   # the arg-min does not depend on the solver: first leaf pixel
    e = ndi.distance_transform_edt(~(lab >= 1))
    # the marched arg-max is a pixel whose exact distance is within a pixel of the exact maximum
    assert e[pmax_f] > e[pmax_e] - 1.0
    exact = O.select_optimal_leaf(lab, dep, P[0, 0], P[0, 2], P[1, 2])
    orig = O.clutter_extrema
    O.clutter_extrema = lambda labels: (pmin_f, pmax_f)
    try:
        marched = O.select_optimal_leaf(lab, dep, P[0, 0], P[0, 2], P[1, 2])
    finally:
        O.clutter_extrema = orig
    assert marched["leaf_id"] == exact["leaf_id"]
