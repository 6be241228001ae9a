"""Oracle (canonical fp32 arithmetic) against oracle/_ref — the reference's OWN OpenCL device
sources compiled for the host (oracle/build_ref.py).  This is what ties the oracle, and
through it the bit-exact CUDA path, to the reference's code rather than to our reading of it.

Bar (north star): distances within 1e-5 relative / 1e-6 * scene-size absolute.  Gradients
are compared away from *tie points*: perpendicular_intersection (common.cl:27-30), the
sharp union (common.cl:60-63) and copysign(1, +-0) pick between candidates by comparing
two distances, so a last-bit difference can swap in a different — equally valid — unit
vector at isolated points (SURVEY.md 7 hard part 3).  Those points are counted and bounded,
and their distances must still agree.
"""
import numpy as np
import pytest

import oracle
from scenes import ALL_NAMES

ref = pytest.importorskip("oracle.ref")
pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")

NAMES = [n for n in ALL_NAMES if n != "cfg_synthetic500"]


def _grid(scene):
    dims = (20, 24, 32) if scene.dimension == 3 else (40, 48, 3)
    corner, step = scene.grid(max(dims))
    return dims, corner, step


@pytest.mark.parametrize("name", NAMES)
def test_grid_eval_matches_compiled_reference(scenes, name):
    s = scenes[name]
    dims, corner, step = _grid(s)
    a = oracle.grid_eval(s.words, corner, step, dims)
    b = ref.grid_eval(s.words, corner, step, dims)
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    dw = np.abs(a[..., 3] - b[..., 3])
    tol = 1e-5 * np.abs(b[..., 3]) + 1e-6 * size
    # Blend values of rounded unions (zero gradient in both, common.cl:52-58) divide by 1 - c^2: where
    # the two normals oppose (deep inside overlapping solids) the rounding of c is amplified without
    # bound, for ANY two implementations of the same formula; the symptom is a "distance" far larger
    # than the scene.  Those points get a 1e-4 relative bar, widened by the square of that excess.
    blend = (np.abs(a[..., :3]).max(axis=-1) == 0) & (np.abs(b[..., :3]).max(axis=-1) == 0)
    excess = np.maximum(1.0, np.abs(b[..., 3]) / (0.25 * size)) ** 2
    tol = np.where(blend, 1e-4 * excess * np.abs(b[..., 3]) + 1e-6 * size, tol)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.all(dw[~np.isnan(dw)] <= tol[~np.isnan(dw)]), "distance off by %g" % np.nanmax(dw - tol)
    dg = np.abs(a[..., :3] - b[..., :3]).max(axis=-1)
    ties = dg > 1e-4
    assert ties.mean() <= 0.02, "%d gradient mismatches" % ties.sum()


def test_synthetic500_sample(scenes):
    s = scenes["cfg_synthetic500"]
    corner, step = s.grid(64)
    a = oracle.grid_eval(s.words, corner, step, (4, 16, 64), x_offset=0)
    b = ref.grid_eval(s.words, corner, step, (4, 16, 64))
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    assert np.all(np.abs(a[..., 3] - b[..., 3]) <= 1e-5 * np.abs(b[..., 3]) + 1e-6 * size)


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith(("cfg", "mp_", "sub_"))])
def test_classification_agrees_where_not_marginal(scenes, name):
    """subdivision_step / mass_properties kernels: identical hit lists and integer sums once
    cells whose distance lies within tolerance of a threshold are set aside."""
    s = scenes[name]
    dims = (16, 16, 16) if s.dimension == 3 else (24, 24, 1)
    corner, step = s.grid(dims[0])
    thr = np.float32(step * np.sqrt(s.dimension) / 2)
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    d = oracle.grid_eval(s.words, corner, step, dims)[..., 3]
    marginal = (np.abs(np.abs(d) - thr) <= 1e-5 * thr + 1e-6 * size)
    a = oracle.subdivision_step(s.words, corner, step, thr, dims)
    b = ref.subdivision_step(s.words, corner, step, thr, dims)
    sa = {tuple(r[:3]) for r in a.tolist() if not marginal[tuple(r[:3])]}
    sb = {tuple(r[:3]) for r in b.tolist() if not marginal[tuple(r[:3])]}
    assert sa == sb
    if s.dimension == 3 and not marginal.any() and not (np.abs(d) <= 1e-6 * size).any():
        sums_a, la = oracle.mass_properties_step(s.words, corner, step, thr, dims)
        sums_b, lb = ref.mass_properties_step(s.words, corner, step, thr, dims)
        assert np.array_equal(sums_a, sums_b)
        assert np.array_equal(la, lb)
