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
    _assert_gradient_mismatches_are_ties(s, a, b, corner, step, dims)


def _assert_gradient_mismatches_are_ties(s, a, b, corner, step, dims, x_offset=0):
    """Where the oracle's direction differs from the reference's, the point must sit on the border
    between two regions of the field (two candidates whose distances tie to rounding): the
    reference's direction then IS the oracle's direction an epsilon away.  A wrong gradient from
    some op would not be found next door."""
    dg = np.abs(a[..., :3] - b[..., :3]).max(axis=-1)
    idx = np.argwhere(dg > 1e-4)
    if len(idx) == 0:
        return
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    pts = (np.asarray(corner, np.float32)[None, :] + np.float32(step) * (idx + [x_offset, 0, 0]).astype(np.float32)).astype(np.float32)
    want = b[tuple(idx.T)][:, :3]
    best = np.full(len(idx), np.inf)
    offsets = [np.array(o, np.float32) for o in
               [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 1), (-1, -1, -1), (1, -1, 0), (-1, 1, 0),
                (0, 1, -1), (0, -1, 1), (1, 0, -1), (-1, 0, 1)]]
    for eps in (2e-6, 2e-5, 2e-4):
        for o in offsets:
            if s.dimension == 2 and o[2] != 0:
                continue
            near = oracle.evaluate_points(s.words, (pts + o * np.float32(eps * size)).astype(np.float32))[:, :3]
            best = np.minimum(best, np.abs(near - want).max(axis=-1))
    assert np.all(best <= 2e-3), "%d of %d gradient mismatches are not ties (worst %.3g)" % (
        int((best > 2e-3).sum()), len(idx), float(best.max()))


def test_config_size_planes(scenes):
    """One full 1024 x 1024 plane of the planetary grid (config C4) and a 2048-long strip of the
    500-box scene at its 2048^3 resolution (C5): the same bars at the sizes the bench runs."""
    for name, n, dims, x0 in (("cfg_planetary", 1024, (1, 1024, 1024), 512), ("cfg_synthetic500", 2048, (2, 6, 2048), 1024)):
        s = scenes[name]
        corner, step = s.grid(n)
        shifted = np.array([np.float32(corner[0] + np.float32(step) * np.float32(x0)), corner[1], corner[2]], np.float32)
        a = oracle.grid_eval(s.words, shifted, step, dims)
        b = ref.grid_eval(s.words, shifted, step, dims)
        size = max(q - p for p, q in zip(s.box_a, s.box_b))
        blend = (np.abs(a[..., :3]).max(axis=-1) == 0) & (np.abs(b[..., :3]).max(axis=-1) == 0)
        excess = np.maximum(1.0, np.abs(b[..., 3]) / (0.25 * size)) ** 2
        tol = np.where(blend, 1e-4 * excess, 1e-5) * np.abs(b[..., 3]) + 1e-6 * size
        assert np.all(np.abs(a[..., 3] - b[..., 3]) <= tol), name
        _assert_gradient_mismatches_are_ties(s, a, b, shifted, step, dims)


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


# ---- the renderers around evaluate(): the oracle's restatements against the reference's own kernels ----
# (rendering/polygon2d.cl, bitmap.cl, matplotlib_slice.cl, ray_caster.cl compiled by oracle/build_ref.py)

needs_renderers = pytest.mark.skipif(not (ref.available() and ref.has_renderers()),
                                     reason="oracle/_ref was built without the rendering kernels")
NAMES_2D = [n for n in NAMES if n.startswith("dsdf2d_") or n == "sub_circle"]
NAMES_3D_SMALL = ["dsdf3d_csg_thing", "dsdf3d_torus", "dsdf3d_box", "dsdf3d_mirror_3d", "cfg_csg_example", "x_gear3d"]


@needs_renderers
@pytest.mark.parametrize("name", NAMES_2D)
@pytest.mark.parametrize("gx", [24, 57])
def test_process_polygon_matches_the_reference_kernel(scenes, name, gx):
    """Same float4 corner grid into both: the integer half (cell types, links, starts) must be
    identical, the vertices (a fixed-count gradient search in fp32) agree to rounding."""
    s = scenes[name]
    corner, step = s.grid(gx)
    corners = oracle.grid_eval(s.words, corner, step, (gx, gx + 3, 1))[:, :, 0, :]
    va, la, sa = oracle.process_polygon(corner, step, corners)
    vb, lb, sb = ref.process_polygon(corner, step, corners)
    assert np.array_equal(la, lb), "links differ"
    assert sorted(sa.tolist()) == sorted(sb.tolist()), "starts of open chains differ"
    crossed = lb != 0xFFFFFFFF
    assert crossed.sum() > 0 or name == "dsdf2d_empty"
    assert np.all(va[~crossed] == 0)
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    assert np.abs(va[crossed] - vb[crossed]).max(initial=0.0) <= 2e-5 * size


@needs_renderers
@pytest.mark.parametrize("name", NAMES_2D[:8])
def test_bitmap_matches_the_reference_kernel(scenes, name):
    from oracle import render
    s = scenes[name]
    size = (96, 64)
    origin, step = render.bitmap_args(s.box_a, s.box_b, size)
    got = render.bitmap(s.words, s.box_a, s.box_b, size)                        # [h][w][3]
    want = ref.bitmap(s.words, origin, step, size).transpose((1, 0, 2))
    differ = np.any(got != want, axis=-1)
    # only pixels whose distance is a rounding error away from zero may fall on the other side
    pts = np.zeros(size + (3,), np.float32)
    pts[..., 0] = (origin[0] + step * np.arange(size[0]))[:, None]
    pts[..., 1] = (origin[1] + step * (size[1] - 1 - np.arange(size[1])))[None, :]
    d = oracle.evaluate_points(s.words, pts.reshape(-1, 3))[:, 3].reshape(size).T
    assert np.all(np.abs(d[differ]) <= 1e-5 * max(q - p for p, q in zip(s.box_a, s.box_b)))


@needs_renderers
@pytest.mark.parametrize("name", NAMES_3D_SMALL + NAMES_2D[:3])
def test_matplotlib_slice_matches_the_reference_kernel(scenes, name):
    s = scenes[name]
    w, h = 40, 28
    corner, step = s.grid(40)
    corner = np.array([corner[0], corner[1], (s.box_a[2] + s.box_b[2]) / 2], np.float32)
    want = ref.matplotlib_slice(s.words, corner, step, (w, h))                   # [h][w][3]
    field = oracle.grid_eval(s.words, corner, step, (w, h, 1))[:, :, 0, :]       # [w][h][4]
    got = np.stack([field[..., 3], field[..., 0], field[..., 1]], axis=-1).transpose((1, 0, 2))
    size = max(q - p for p, q in zip(s.box_a, s.box_b))
    assert np.all(np.abs(got[..., 0] - want[..., 0]) <= 1e-5 * np.abs(want[..., 0]) + 1e-6 * size)
    ties = np.abs(got[..., 1:] - want[..., 1:]).max(axis=-1) > 1e-4
    assert ties.mean() <= 0.02


@needs_renderers
@pytest.mark.parametrize("name", NAMES_3D_SMALL)
@pytest.mark.parametrize("options", [0, 1, 2])
def test_ray_caster_matches_the_reference_kernel(scenes, name, options):
    """Plain, false-colour and zebra pictures: the oracle's restatement of ray_caster.cl against the
    kernel itself (same arguments).  Rays that graze a surface amplify last-bit differences of
    evaluate() into different step counts, so a small share of pixels may differ visibly."""
    from oracle import render
    s = scenes[name]
    size = (120, 90)
    args = render.ray_cast_args(s.box_a, s.box_b, size)
    got = render.ray_cast(s.words, s.box_a, s.box_b, size, options=options)     # [h][w][3]
    want = ref.ray_caster(s.words, *args, options, size).transpose((1, 0, 2))
    keep = np.ones(got.shape[:2], bool)
    if options & 1:
        # False colour shows step counts.  For a ray that MISSES, the kernel goes on to trace a light ray
        # from a point at infinity: whether evaluate() yields inf or NaN there (quaternion vs matrix form
        # of the same transformation) decides whether that loop stops at once or runs its 100 steps.
        # Undefined input, not compared; the pixels must be background in the plain picture.
        plain = ref.ray_caster(s.words, *args, 0, size).transpose((1, 0, 2)).astype(int)
        background = (plain[..., 0] == plain[..., 1]) & (plain[..., 2] >= plain[..., 0])
        miss = (want[..., 0].astype(int) - got[..., 0].astype(int) >= 60) & (got[..., 0] <= 60)
        assert np.all(background[miss])
        keep = ~miss
        assert keep.mean() > 0.2
    err = (got.astype(np.float32) / 255 - want.astype(np.float32) / 255)[keep]
    assert float(np.mean(err * err)) <= 1e-3                                     # tests/tools.py:64-79, the reference's own bar
    assert np.mean(np.abs(got.astype(int) - want.astype(int)).max(axis=-1)[keep] > 8) <= 0.01
