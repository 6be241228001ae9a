"""Pin the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8(c)); no GPU involved.

  reference tests/test_mass_properties.py:16-108   analytic volume / centroid / inertia
  reference tests/test_subdivision.py:110-161      leaf-block corner sets
  reference tests/test_dsdf.py:113-192             gradient / distance properties, 32 shapes
"""
import itertools
import math

import numpy as np
import pytest

import oracle
from oracle import host
from scenes import ALL_NAMES

drunk = None


def _quat_matrix(axis, angle_deg):
    ax = np.array(axis, float)
    ax /= np.linalg.norm(ax)
    phi = math.radians(angle_deg) / 2
    v, w = ax * math.sin(phi), math.cos(phi)
    x, y, z = v
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


_R = _quat_matrix((7, 11, 13), 17)
ANALYTIC = {
    "mp_unit_box": (1.0, (0, 0, 0), np.identity(3) * 2 / 12),
    "mp_cylinder": (math.pi * 32, (0, 0, 1), np.diag([(3 * 16 + 4) / 6, (3 * 16 + 4) / 6, 16.0]) * math.pi * 16),
    "mp_sphere": (4 * math.pi / 3, (0, 0, 0), np.identity(3) * (4 * math.pi / 3) * 2 / 5),
    "mp_two_boxes": (16.0, (0, 0, 0), None),
    "mp_hemisphere": (2 * math.pi * 8 / 3, (0, -6 / 8, 0), None),
    "mp_translated_sphere": (4 * math.pi / 3, (10, 11, 7), None),
    "mp_translated_and_rotated_hemisphere": (2 * math.pi * 8 / 3, (2, 0, -6 / 8), None),
    "mp_not_hammer": (96.0, (0, 0, 0), np.diag([1120.0, 1120.0, 192.0])),
    "mp_drunk_box": (30.0, (0, 0, 0), _R @ (np.diag([9 + 25, 4 + 25, 4 + 9]) * 30 / 12) @ _R.T),
}


@pytest.mark.parametrize("name", sorted(ANALYTIC))
def test_mass_properties_analytic(scenes, name):
    """reference tests/test_mass_properties.py:99-108, same resolution and tolerances"""
    volume, centroid, inertia = ANALYTIC[name]
    s = scenes[name]
    precision = 2e-3
    vol, cen, it = host.mass_properties(s.words, s.box_a, s.box_b, 10 * precision, 64)
    assert vol == pytest.approx(volume, abs=1e-4, rel=precision)
    assert tuple(cen) == pytest.approx(centroid, abs=1e-4, rel=precision)
    if inertia is not None:
        assert np.allclose(it, inertia, rtol=precision, atol=precision * np.abs(inertia).max() * 1e-1)


def test_block_corners_cube(scenes):
    """reference tests/test_subdivision.py:110-127"""
    s = scenes["sub_box10"]
    _, blocks = host.subdivision(s.words, s.box_a, s.box_b, 3, 1, True, 4)
    assert blocks[0][2] == 1 and blocks[0][4] == 1
    corners = {tuple(b[1]) for b in blocks}
    expected = set(itertools.product([-5.5, -2.5, 0.5, 3.5], repeat=3)) - set(itertools.product([-2.5, 0.5], repeat=3))
    assert corners == expected


def test_block_corners_circle(scenes):
    """reference tests/test_subdivision.py:130-161"""
    resolution, grid_size = 0.1, 8
    step = resolution * (grid_size - 1)
    radius = (grid_size * step - resolution) / 2
    threshold = math.sqrt(2) * step / 2
    s = scenes["sub_circle"]
    _, blocks = host.subdivision(s.words, s.box_a, s.box_b, 2, resolution, True, grid_size)
    assert blocks[0][2] == resolution and blocks[0][4] == 1
    r = [-radius - 0.5 * resolution + i * step for i in range(grid_size)]
    expected = set()
    for cx, cy in itertools.product(r, repeat=2):
        if radius - threshold < math.hypot(cx + step / 2, cy + step / 2) < radius + threshold:
            expected.add((cx, cy))
    got = {(b[1][0], b[1][1]) for b in blocks}
    assert len(got) == len(expected)
    for g in got:
        assert any(abs(g[0] - e[0]) < 1e-9 and abs(g[1] - e[1]) < 1e-9 for e in expected)


# ---- reference tests/test_dsdf.py on the shapes of reference tests/data.py:53-96 ----------------

DSDF = [n for n in ALL_NAMES if n.startswith("dsdf")]


def _dsdf_grid(scene):
    size = (16, 16, 16) if scene.dimension == 3 else (16, 16, 3)
    corner = np.array([-s / 2 for s in size], np.float32)
    return size, corner, np.float32(1)


@pytest.mark.parametrize("name", [n for n in DSDF if n.startswith("dsdf2d")])
def test_2d_direction_has_zero_z(scenes, name):
    """test_dsdf.py:113-118 — exactly zero"""
    size, corner, step = _dsdf_grid(scenes[name])
    v = oracle.grid_eval(scenes[name].words, corner, step, size)
    assert np.all(v[..., 2] == 0)


@pytest.mark.parametrize("name", DSDF)
def test_direction_unit_length(scenes, name):
    """test_dsdf.py:121-126 (pytest.approx(1): rel 1e-6); the rounded-union blend region
    returns a zero gradient by design (common.cl:56 TODO) and no test shape uses it"""
    size, corner, step = _dsdf_grid(scenes[name])
    v = oracle.grid_eval(scenes[name].words, corner, step, size)
    n2 = (v[..., :3].astype(np.float64) ** 2).sum(-1)
    assert np.allclose(n2, 1.0, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("name", DSDF)
def test_distance_is_lower_bound(scenes, name):
    """test_dsdf.py:154-160 with actual_distance_to_surface of test_dsdf.cl:49-72: the
    distance to the nearest grid point of opposite sign bounds the returned distance"""
    size, corner, step = _dsdf_grid(scenes[name])
    v = oracle.grid_eval(scenes[name].words, corner, step, size)[..., 3]
    idx = np.stack(np.meshgrid(*[np.arange(s) for s in size], indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    sign = np.sign(v).reshape(-1)
    d = v.reshape(-1)
    for sg in (-1.0, 0.0, 1.0):
        mine = np.nonzero(sign == sg)[0]
        other = np.nonzero(sign != sg)[0]
        if len(mine) == 0:
            continue
        if len(other) == 0:
            continue  # MAXFLOAT in the reference: always satisfied
        diff = idx[mine][:, None, :] - idx[other][None, :, :]
        nearest = np.sqrt((diff * diff).sum(-1).min(1)) * float(step)
        assert np.all(d[mine] <= nearest + 1e-5)


@pytest.mark.parametrize("name", DSDF)
def test_direction_matches_finite_differences(scenes, name):
    """test_dsdf.py:163-192 with estimate_direction of test_dsdf.cl:7-30"""
    s = scenes[name]
    size, corner, step = _dsdf_grid(s)
    eps = np.float32(0.05)
    pts = np.stack(np.meshgrid(*[corner[i] + step * np.arange(size[i], dtype=np.float32) for i in range(3)],
                               indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    center = oracle.evaluate_points(s.words, pts)
    plus = np.stack([oracle.evaluate_points(s.words, pts - eps * np.eye(3, dtype=np.float32)[i])[:, 3] for i in range(3)], -1)
    minus = np.stack([oracle.evaluate_points(s.words, pts + eps * np.eye(3, dtype=np.float32)[i])[:, 3] for i in range(3)], -1)
    fd1 = center[:, 3:4] - plus
    fd2 = minus - center[:, 3:4]
    smooth = np.linalg.norm(fd1 - fd2, axis=1) <= 1e-4
    fd = (fd1 + fd2) / 2
    norm = np.linalg.norm(fd, axis=1)
    ok = smooth & (norm > 0)
    fd = fd[ok] / norm[ok][:, None]
    assert np.all(np.linalg.norm(center[ok, :3] - fd, axis=1) < 1e-2)


@pytest.mark.parametrize("name", ["cfg_airfoil", "dsdf2d_polygon2d_non_convex", "dsdf2d_polygon2d_collinear_consecutive_edges",
                                  "dsdf2d_polygon2d_collinear_non_consecutive_edges", "dsdf2d_polygon2d_parallel_same_direction_edges",
                                  "dsdf2d_polygon2d_square", "dsdf2d_polygon2d_triangle", "dsdf2d_nonconvex_shell1"])
def test_polygon2d_product_formulation_is_bit_identical(scenes, name):
    """The CUDA path evaluates polygon2d with a branch-free loop (running minimum + edge index,
    clamped t, rolling crossing test; csrc/cc_ops.cuh cc_polygon2d_v).  Its restatement in C must give
    the reference formulation's bits everywhere: dense random points, points on and next to every
    vertex and edge, and special operands."""
    s = scenes[name]
    rng = np.random.default_rng(11)
    a, b = np.asarray(s.box_a, np.float64), np.asarray(s.box_b, np.float64)
    a, b = np.where(np.isfinite(a), a, -10.0), np.where(np.isfinite(b), b, 10.0)
    n = 400_000
    pts = (a + (b - a) * (rng.random((n, 3)) * 1.6 - 0.3)).astype(np.float32)
    # lattice points: exactly on vertices / edges of the integer-coordinate test polygons, and one ulp off
    k = np.arange(-2, 9, dtype=np.float32) * 0.5
    gx, gy = np.meshgrid(k, k, indexing="ij")
    lattice = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size, np.float32)], -1)
    near = np.concatenate([lattice, np.nextafter(lattice, np.float32(9)), np.nextafter(lattice, np.float32(-9))])
    special = np.array([[np.nan, 0, 0], [0, np.nan, 0], [np.inf, 1, 0], [1, -np.inf, 0], [3e38, -3e38, 0],
                        [1e-45, 0, 0], [-0.0, 0.0, 0], [1e-39, 1e-39, 0]], np.float32)
    pts = np.concatenate([pts, near, special])
    if s.dimension == 2:
        pts[:, 2] = 0
    L = oracle.lib()
    try:
        want = oracle.evaluate_points(s.words, pts)
        L.oracle_set_polygon_formulation(1)
        got = oracle.evaluate_points(s.words, pts)
    finally:
        L.oracle_set_polygon_formulation(0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
