"""GPU parity: every kernel of the CUDA path against the CPU oracle on the same program
words and the same fp32 inputs.  The canonical arithmetic (DESIGN.md "cc-arith") is
mirrored exactly by the oracle, so the bar is BIT-EXACT for distances, gradients, index
lists and integer sums (tolerance 0; NaNs must coincide)."""
import ctypes

import numpy as np
import pytest

import oracle
from scenes import ALL_NAMES

pytestmark = pytest.mark.gpu

SMALL = [n for n in ALL_NAMES if n != "cfg_synthetic500"]


def _same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def _f4(arr):
    return np.stack([arr["x"], arr["y"], arr["z"], arr["w"]], axis=-1)


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    # interpreter tier only, so that every test below knows which kernels it exercises; the
    # specialised tier is switched on explicitly by the tests at the end of this file
    old = _lib.check(_lib.lib().cc_set_jit_mode(0))
    yield codecad_b200
    _lib.check(_lib.lib().cc_set_tuning(0, 0))
    _lib.check(_lib.lib().cc_set_jit_mode(old))


def _dims_for(scene):
    return (20, 24, 32) if scene.dimension == 3 else (40, 48, 3)


@pytest.mark.parametrize("name", ALL_NAMES)
def test_grid_eval_bit_exact(cb, scenes, name):
    s = scenes[name]
    dims = _dims_for(s) if name != "cfg_synthetic500" else (8, 12, 32)
    corner, step = s.grid(max(dims))
    want = oracle.grid_eval(s.words, corner, step, dims)
    got = _f4(cb.grid_eval(s.compiled(), corner, step, dims))
    assert got.shape == want.shape
    assert _same(got, want), "max |diff| %g" % np.nanmax(np.abs(got - want))


@pytest.mark.parametrize("pts", [1, 2, 4])
@pytest.mark.parametrize("space", [1, 2])
@pytest.mark.parametrize("name", ["cfg_planetary", "cfg_airfoil", "cfg_menger_sponge", "cfg_synthetic32",
                                  "dsdf3d_extreme_twisted_revolve", "dsdf2d_gear"])
def test_grid_eval_all_kernel_variants(cb, scenes, name, pts, space):
    from codecad_b200 import _lib
    s = scenes[name]
    dims = (16, 9, 37) if s.dimension == 3 else (33, 17, 2)  # ragged: exercises tile tails
    corner, step = s.grid(40)
    want = oracle.grid_eval(s.words, corner, step, dims)
    _lib.check(_lib.lib().cc_set_tuning(pts, space))
    try:
        got = _f4(cb.grid_eval(s.compiled(), corner, step, dims))
    finally:
        _lib.check(_lib.lib().cc_set_tuning(0, 0))
    assert _same(got, want)


def test_grid_eval_slab_offset_matches_unsharded(cb, scenes):
    s = scenes["cfg_csg_example"]
    dims = (32, 16, 16)
    corner, step = s.grid(32)
    full = _f4(cb.grid_eval(s.compiled(), corner, step, dims))
    for world in (2, 3):
        parts = []
        for rank in range(world):
            x0, x1 = cb.grid_eval.__globals__["slab_range"](dims[0], rank, world)
            parts.append(_f4(cb.grid_eval(s.compiled(), corner, step, (x1 - x0, dims[1], dims[2]), x_offset=x0)))
        assert _same(np.concatenate(parts, axis=0), full)


@pytest.mark.parametrize("name", ["cfg_csg_example", "cfg_planetary", "dsdf3d_torus", "dsdf2d_mirror_2d"])
def test_grid_eval_pymcubes_layout(cb, scenes, name):
    s = scenes[name]
    dims = (12, 10, 16) if s.dimension == 3 else (12, 10, 1)
    corner, step = s.grid(16)
    want = oracle.grid_eval_pymcubes(s.words, corner, step, dims)
    got = cb.grid_eval_pymcubes(s.compiled(), corner, step, dims)
    assert _same(got, want)
    # and it is the y-flipped, axis-swapped view of the float4 grid's distances
    f4 = _f4(cb.grid_eval(s.compiled(), corner, step, dims))[..., 3]
    assert _same(got, np.transpose(f4, (1, 0, 2))[::-1])


def _step_buffers(cb, n):
    from codecad_b200.cl_util import Buffer
    from codecad_b200.geometry import UCHAR4
    counter = Buffer(np.uint32, 1)
    lst = Buffer(UCHAR4, n)
    counter.enqueue_zero_fill_compatible()
    return counter, lst


@pytest.mark.parametrize("name", SMALL)
def test_subdivision_step_matches_oracle(cb, scenes, name):
    from codecad_b200.cl_util import opencl_manager
    from codecad_b200.geometry import Vector
    s = scenes[name]
    dims = (16, 16, 16) if s.dimension == 3 else (24, 24, 1)
    corner, step = s.grid(16 if s.dimension == 3 else 24)
    thr = np.float32(step * np.sqrt(s.dimension) / 2)
    want = oracle.subdivision_step(s.words, corner, step, thr, dims)
    counter, lst = _step_buffers(cb, dims[0] * dims[1] * dims[2])
    prog = s.compiled().program_buffer()
    ev = opencl_manager.k.subdivision_step(dims, None, prog, Vector(*corner).as_float4(), step, thr, counter, lst)
    n = int(counter.read(wait_for=[ev])[0])
    got = lst.read()[:n]
    got = np.stack([got["x"], got["y"], got["z"], got["w"]], axis=-1)
    assert n == len(want)
    assert _same(got, want)  # same cells AND same (INDEX3) order


@pytest.mark.parametrize("thr_scale", [0.0, 1.0])
@pytest.mark.parametrize("name", [n for n in SMALL if n.startswith(("cfg", "mp_", "dsdf3d"))])
def test_mass_properties_step_matches_oracle(cb, scenes, name, thr_scale):
    from codecad_b200.cl_util import Buffer, opencl_manager
    from codecad_b200.geometry import Vector
    s = scenes[name]
    dims = (16, 12, 20)
    corner, step = s.grid(20)
    thr = np.float32(thr_scale * step * np.sqrt(3) / 2)
    want_sums, want_list = oracle.mass_properties_step(s.words, corner, step, thr, dims)
    counter, lst = _step_buffers(cb, dims[0] * dims[1] * dims[2])
    sums = Buffer(np.uint32, 10)
    sums.enqueue_zero_fill_compatible()
    prog = s.compiled().program_buffer()
    ev = opencl_manager.k.mass_properties(dims, None, prog, Vector(*corner).as_float4(), step, thr, sums, counter, lst)
    n = int(counter.read(wait_for=[ev])[0])
    got = lst.read()[:n]
    got = np.stack([got["x"], got["y"], got["z"], got["w"]], axis=-1)
    assert _same(sums.read(), want_sums)
    assert n == len(want_list) and _same(got, want_list)


# ---- whole-hierarchy drivers against the oracle's restatement of the reference host loops ----

SUBDIV_CASES = [
    ("sub_box10", 1.0, 4, True),          # reference tests/test_subdivision.py:110-127
    ("sub_circle", 0.1, 8, True),         # :130-161
    ("cfg_csg_example", 100 / 512, 16, True),
    ("cfg_csg_example", 100 / 128, 8, False),
    ("cfg_menger_sponge", 0.75, 16, True),
    ("dsdf2d_gear", 0.05, 8, True),
    ("dsdf3d_mirror_3d", 0.06, 6, True),
]


@pytest.mark.parametrize("name,resolution,grid,overlap", SUBDIV_CASES)
def test_subdivision_matches_oracle(cb, scenes, name, resolution, grid, overlap):
    from oracle import host
    s = scenes[name]
    want_dims, want = host.subdivision(s.words, s.box_a, s.box_b, s.dimension, resolution, overlap, grid)
    _, got_dims, got = cb.subdivision(s.compiled(), resolution, overlap, grid)
    assert tuple(got_dims) == tuple(want_dims)
    key = lambda b: tuple(b[3])
    want = sorted(want, key=key)
    got = sorted(got, key=key)
    assert [tuple(b[3]) for b in got] == [tuple(b[3]) for b in want]       # int corners, bit-exact
    assert [tuple(b[1]) for b in got] == [tuple(b[1]) for b in want]       # float64 corners, bit-exact
    assert all(g[2] == w[2] and g[4] == w[4] and tuple(g[0]) == tuple(w[0]) for g, w in zip(got, want))
    # sharded over 3 ranks: disjoint union equals the whole
    parts = []
    for rank in range(3):
        parts += cb.subdivision(s.compiled(), resolution, overlap, grid, rank=rank, world=3)[2]
    assert sorted(tuple(b[3]) for b in parts) == [tuple(b[3]) for b in want]


MASS_CASES = [
    ("cfg_airfoil", 1.0, 64), ("cfg_airfoil", 0.5, 64), ("cfg_csg_example", 1.0, 16),
    ("cfg_menger_sponge", 0.8, 32), ("cfg_planetary", 0.8, 64), ("cfg_synthetic32", 1.5, 32),
    ("mp_unit_box", 0.02, 64), ("mp_drunk_box", 0.05, 8), ("x_gear3d", 0.1, 16),
]


@pytest.mark.parametrize("name,resolution,grid", MASS_CASES)
def test_mass_properties_matches_oracle(cb, scenes, name, resolution, grid):
    from oracle import host
    s = scenes[name]
    w_vol, w_cen, w_inertia = host.mass_properties(s.words, s.box_a, s.box_b, resolution, grid)
    got = cb.mass_properties(s.compiled(), resolution, grid)
    # north-star bar: 1e-6 relative.  Cell classification is bit-exact, so the only
    # difference left is float64 summation order (~1e-15).
    assert got.volume == pytest.approx(w_vol, rel=1e-12)
    assert tuple(got.centroid) == pytest.approx(tuple(w_cen), rel=1e-9, abs=1e-9 * max(1.0, abs(w_vol)) ** (1 / 3))
    scale = np.abs(w_inertia).max()
    assert np.allclose(got.inertia_tensor, w_inertia, rtol=1e-9, atol=1e-9 * scale)


# reference tests/test_mass_properties.py:16-108, run through the CUDA path
ANALYTIC = {
    "mp_unit_box": (1.0, (0, 0, 0), np.identity(3) * 2 / 12),
    "mp_cylinder": (np.pi * 32, (0, 0, 1), np.diag([(3 * 16 + 4) / 6, (3 * 16 + 4) / 6, 16.0]) * np.pi * 16),
    "mp_sphere": (4 * np.pi / 3, (0, 0, 0), np.identity(3) * (4 * np.pi / 3) * 2 / 5),
    "mp_two_boxes": (16.0, (0, 0, 0), None),
    "mp_hemisphere": (2 * np.pi * 8 / 3, (0, -6 / 8, 0), None),
    "mp_translated_sphere": (4 * np.pi / 3, (10, 11, 7), None),
    "mp_translated_and_rotated_hemisphere": (2 * np.pi * 8 / 3, (2, 0, -6 / 8), None),
    "mp_not_hammer": (96.0, (0, 0, 0), np.diag([1120.0, 1120.0, 192.0])),
}


@pytest.mark.parametrize("name", sorted(ANALYTIC))
def test_mass_properties_known_answers(cb, scenes, name):
    volume, centroid, inertia = ANALYTIC[name]
    precision = 2e-3
    got = cb.mass_properties(scenes[name].compiled(), 10 * precision)
    assert got.volume == pytest.approx(volume, abs=1e-4, rel=precision)
    assert tuple(got.centroid) == pytest.approx(centroid, abs=1e-4, rel=precision)
    if inertia is not None:
        assert np.allclose(got.inertia_tensor, inertia, rtol=precision)


# ---- scene-specialised (NVRTC) kernels: same op library, same body -> same bits -----------------

JIT_SCENES = ["cfg_synthetic500", "cfg_csg_example", "cfg_menger_sponge", "cfg_airfoil", "cfg_planetary", "cfg_synthetic32",
              "dsdf3d_extreme_twisted_revolve", "dsdf2d_gear", "dsdf2d_regular_polygon3", "dsdf3d_rotated_pattern_3d",
              "dsdf3d_revolved_pentagon", "x_repetition", "x_smooth_isect", "dsdf2d_polygon2d_non_convex"]


@pytest.mark.parametrize("pts", [1, 2])
@pytest.mark.parametrize("name", JIT_SCENES)
def test_specialized_grid_eval_bit_exact(cb, scenes, name, pts):
    from codecad_b200.cl_util.buffer import ProgramBuffer
    s = scenes[name]
    prog = ProgramBuffer(s.words)
    secs = prog.specialize(pts, ProgramBuffer.SINK_FLOAT4 | ProgramBuffer.SINK_PYMCUBES)
    assert secs > 0 and prog.use_specialized(True)
    dims = (16, 9, 37) if s.dimension == 3 else (33, 17, 2)
    corner, step = s.grid(40)
    want = oracle.grid_eval(s.words, corner, step, dims)
    got = _f4(cb.grid_eval(prog, corner, step, dims))
    assert _same(got, want)
    assert _same(cb.grid_eval_pymcubes(prog, corner, step, dims), oracle.grid_eval_pymcubes(s.words, corner, step, dims))
    # switching back to the interpreter gives the same bits again
    assert not prog.use_specialized(False)
    assert _same(_f4(cb.grid_eval(prog, corner, step, dims)), want)


@pytest.mark.parametrize("name", ["cfg_airfoil", "cfg_csg_example", "cfg_planetary"])
def test_specialized_hierarchy_matches_oracle(cb, scenes, name):
    """subdivision() and mass_properties() through specialised classify / mass kernels."""
    from oracle import host
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200 import CompiledScene
    s = scenes[name]
    scene = s.compiled()
    scene._buffer = ProgramBuffer(s.words)
    scene._buffer.specialize(2, ProgramBuffer.SINK_CLASSIFY | ProgramBuffer.SINK_MASS)
    res = {"cfg_airfoil": 1.0, "cfg_csg_example": 1.0, "cfg_planetary": 0.8}[name]
    w_vol, w_cen, w_inertia = host.mass_properties(s.words, s.box_a, s.box_b, res, 32)
    got = cb.mass_properties(scene, res, 32)
    assert got.volume == pytest.approx(w_vol, rel=1e-12)
    assert np.allclose(got.inertia_tensor, w_inertia, rtol=1e-9, atol=1e-9 * np.abs(w_inertia).max())
    _, want = host.subdivision(s.words, s.box_a, s.box_b, s.dimension, res, True, 8)
    got_blocks = cb.subdivision(scene, res, True, 8)[2]
    assert sorted(tuple(b[3]) for b in got_blocks) == sorted(tuple(b[3]) for b in want)


@pytest.mark.parametrize("pts", [4])
@pytest.mark.parametrize("name", ["cfg_planetary", "cfg_synthetic32", "dsdf2d_gear"])
def test_specialized_four_points_per_thread(cb, scenes, name, pts):
    from codecad_b200.cl_util.buffer import ProgramBuffer
    s = scenes[name]
    prog = ProgramBuffer(s.words)
    prog.specialize(pts, ProgramBuffer.SINK_FLOAT4)
    dims = (16, 9, 37) if s.dimension == 3 else (33, 17, 2)
    corner, step = s.grid(40)
    assert _same(_f4(cb.grid_eval(prog, corner, step, dims)), oracle.grid_eval(s.words, corner, step, dims))


def test_tiered_execution_modes(cb, scenes):
    """Default behaviour: interpreter first, specialised kernel takes over when its background
    compile has finished; mode 2 waits at first use.  Same bits in every tier."""
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    L = _lib.lib()
    s = scenes["cfg_menger_sponge"]
    dims = (16, 9, 37)
    corner, step = s.grid(40)
    want = oracle.grid_eval(s.words, corner, step, dims)
    try:
        assert _lib.check(L.cc_set_jit_mode(1)) == 0
        prog = ProgramBuffer(s.words)
        assert _same(_f4(cb.grid_eval(prog, corner, step, dims)), want)      # interpreter or specialised
        n, secs = prog.wait_specialized(ProgramBuffer.SINK_FLOAT4)
        assert n == 2 and secs >= 0      # the plain float4 kernel and, for this two-part scene, the part-culling pair
        assert prog.use_specialized(True)
        launches0, _ = _lib.counters()
        assert _same(_f4(cb.grid_eval(prog, corner, step, dims)), want)      # specialised now
        assert _lib.counters()[0] == launches0 + 2                            # brick-centre pass + one CTA per brick
        # a second program with the same words hits the in-memory cubin cache
        assert _lib.check(L.cc_set_jit_mode(2)) == 1
        prog2 = ProgramBuffer(s.words)
        assert _same(_f4(cb.grid_eval(prog2, corner, step, dims)), want)
        assert prog2.use_specialized(True)
        # the hierarchy drivers go through the same switch
        a = scenes["cfg_airfoil"]
        from oracle import host
        vol, _, _ = host.mass_properties(a.words, a.box_a, a.box_b, 1.0, 32)
        assert cb.mass_properties(a.compiled(), 1.0, 32).volume == pytest.approx(vol, rel=1e-12)
    finally:
        _lib.check(L.cc_set_jit_mode(0))


@pytest.mark.parametrize("seg", ["7", "40"])
@pytest.mark.parametrize("name", ["cfg_planetary", "cfg_menger_sponge", "cfg_synthetic32", "x_smooth_isect"])
def test_specialized_segmented_programs(cb, scenes, name, seg, monkeypatch):
    """Large programs are generated as __noinline__ segments with table-driven primitives; forcing
    tiny segments on ordinary scenes exercises every cross-segment value path."""
    from codecad_b200.cl_util.buffer import ProgramBuffer
    monkeypatch.setenv("CODECAD_B200_JIT_SEGMENT_OPS", seg)
    s = scenes[name]
    prog = ProgramBuffer(s.words)
    prog.specialize(2, ProgramBuffer.SINK_FLOAT4 | ProgramBuffer.SINK_CLASSIFY)
    dims = (16, 9, 37) if s.dimension == 3 else (33, 17, 2)
    corner, step = s.grid(40)
    assert _same(_f4(cb.grid_eval(prog, corner, step, dims)), oracle.grid_eval(s.words, corner, step, dims))
    from codecad_b200.cl_util import opencl_manager
    from codecad_b200.geometry import Vector
    sdims = (17, 13, 29) if s.dimension == 3 else (40, 37, 1)
    thr = np.float32(step * np.sqrt(s.dimension) / 2)
    want = oracle.subdivision_step(s.words, corner, step, thr, sdims)
    counter, lst = _step_buffers(cb, sdims[0] * sdims[1] * sdims[2])
    ev = opencl_manager.k.subdivision_step(sdims, None, prog, Vector(*corner).as_float4(), step, thr, counter, lst)
    n = int(counter.read(wait_for=[ev])[0])
    got = lst.read()[:n]
    got = np.stack([got["x"], got["y"], got["z"], got["w"]], axis=-1)
    assert n == len(want) and _same(got, want)


# ---- evaluate() at arbitrary points (cc_evaluate_points) ----------------------------------------

@pytest.mark.parametrize("tier", ["interpreter", "specialised"])
@pytest.mark.parametrize("name", ["cfg_planetary", "cfg_csg_example", "dsdf2d_gear", "dsdf3d_extreme_twisted_revolve",
                                  "x_repetition", "cfg_synthetic32", "cfg_airfoil", "dsdf2d_polygon2d_non_convex",
                                  "dsdf2d_polygon2d_collinear_consecutive_edges"])
def test_evaluate_points_bit_exact(cb, scenes, name, tier):
    from codecad_b200.cl_util.buffer import ProgramBuffer
    s = scenes[name]
    rng = np.random.default_rng(7)
    a, b = np.asarray(s.box_a, np.float64), np.asarray(s.box_b, np.float64)
    a, b = np.where(np.isfinite(a), a, -10.0), np.where(np.isfinite(b), b, 10.0)
    n = 1000 + 37                                           # ragged tail tile
    pts = (a + (b - a) * (rng.random((n, 3)) * 1.4 - 0.2)).astype(np.float32)
    if s.dimension == 2:
        pts[:, 2] = 0
    pts[:6] = [[0, 0, 0], [np.inf, 0, 0], [np.nan, 1, 2], [1, -np.inf, 0], [3e38, 3e38, 0], [1e-40, -1e-42, 0]]
    if s.words.size > 8:                                    # and points exactly on polygon vertices / edges
        pts[6:10] = [[0, 0, 0], [3, 3, 0], [1.5, 0, 0], [4, 0, 0]]
    prog = ProgramBuffer(s.words)
    if tier == "specialised":
        prog.specialize(2, ProgramBuffer.SINK_POINTS)
        assert prog.use_specialized(True)
    want = oracle.evaluate_points(s.words, pts)
    got = cb.evaluate_points(prog, pts)
    assert got.shape == (n, 4) and _same(got, want)
    assert cb.evaluate_points(prog, np.zeros((0, 3), np.float32)).shape == (0, 4)
