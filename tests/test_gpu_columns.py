"""Columns: dense float4 grids evaluate the micro-ops that cannot see the grid's z — the
2-D profile under an extrusion — once per z-column.  Must not change a bit: against the CPU oracle on
windows of several scales, against the same library with the column kernel switched off on larger grids,
and on grids with cell centres exactly on the coordinate planes, where a transform whose z coefficient is
rounding residue does depend on z and the per-column check has to send the column down the full path."""
import numpy as np
import pytest

import oracle
from scenes import ALL_NAMES, COLUMN_NAMES

pytestmark = pytest.mark.gpu
# the configs and 2-D fixtures with column-invariant work, and tests/golden/column_scenes.npz: extrusions along x, y, z and
# oblique, under mirror / symmetry / offset / shell / nested transformations, half turns, mixed with solids
# (the ones the loader finds nothing in run too: they must take the other kernels and still agree)
COLUMN_SCENES = ["cfg_planetary", "cfg_airfoil", "x_gear3d", "dsdf2d_gear", "dsdf2d_mirror_2d", "dsdf2d_rotated_pattern_2d",
                 "dsdf2d_bin_counter_11"] + COLUMN_NAMES


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    L = _lib.lib()
    old_jit = _lib.check(L.cc_set_jit_mode(2))     # compile at first use and wait: the column kernel is an NVRTC kernel
    yield codecad_b200
    _lib.check(L.cc_set_columns_mode(1))
    _lib.check(L.cc_set_jit_mode(old_jit))


def _f4(arr):
    return np.stack([arr["x"], arr["y"], arr["z"], arr["w"]], axis=-1)


def _box(s):
    a, b = np.array(s.box_a, dtype=np.float64), np.array(s.box_b, dtype=np.float64)
    if s.dimension == 2:
        a[2], b[2] = -1.0, 1.0
    return a, b


def test_which_scenes_have_columns(scenes):
    from codecad_b200 import _lib
    have = sorted(n for n in ALL_NAMES if _lib.decode_program(scenes[n].words)[0].column_invariant_percent)
    none = {"col_star_oblique", "col_repeated", "colr_10", "colr_23"}
    assert set(COLUMN_SCENES) - none <= set(have)
    assert not none & set(have)     # an oblique extrusion axis; a repetition (reads every axis); solids in general position
    assert _lib.decode_program(scenes["cfg_planetary"].words)[0].column_invariant_percent >= 40
    assert _lib.decode_program(scenes["cfg_csg_example"].words)[0].column_invariant_percent == 0
    assert _lib.decode_program(scenes["cfg_planetary"].words)[0].column_axis == 2       # gears extruded along z
    assert _lib.decode_program(scenes["cfg_airfoil"].words)[0].column_axis != 2         # the wing's span is not the grid's z


@pytest.mark.parametrize("name", COLUMN_SCENES)
def test_columns_bit_exact_vs_oracle(cb, scenes, name):
    from codecad_b200 import _lib
    _lib.check(_lib.lib().cc_set_columns_mode(1))
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(11)
    a, b = _box(s)
    size = float(max(b - a))
    corner, step = s.grid(48)
    windows = [(corner, step, (48, 40, 33))]
    for frac, dims in ((0.3, (40, 33, 48)), (0.05, (33, 48, 40)), (0.01, (24, 40, 64))):
        for _ in range(2):
            centre = a + (b - a) * rng.uniform(0.2, 0.8, 3)
            st = np.float32(size * frac / 32)
            windows.append(((centre - st * np.array(dims) / 2).astype(np.float32), st, dims))
    # cell centres exactly on x = 0, y = 0 and z = 0 (a power-of-two step, the corner a multiple of it)
    st = np.float32(2.0 ** np.floor(np.log2(size / 40)))
    windows.append((np.array([-20 * st, -16 * st, -24 * st], np.float32), st, (40, 32, 48)))
    launches0 = _lib.counters()[0]
    for corner, step, dims in windows:
        want = oracle.grid_eval(s.words, corner, step, dims)
        got = _f4(cb.grid_eval(scene, corner, step, dims))
        assert np.array_equal(got, want, equal_nan=True), "%s step %g: %d values differ" % (name, step, int((got != want).sum()))
    assert _lib.counters()[0] - launches0 >= len(windows)
    if _lib.decode_program(s.words)[0].column_invariant_percent and name != "cfg_planetary":
        assert _lib.counters()[0] - launches0 >= 2 * len(windows)      # column pass + brick kernel at least


@pytest.mark.parametrize("name", COLUMN_SCENES)
def test_columns_equal_the_other_kernels(cb, scenes, name):
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(23)
    a, b = _box(s)
    size = float(max(b - a))
    cases = [(1.05, (128, 128, 128), 0, None), (0.3, (130, 70, 90), 5, None), (0.04, (64, 136, 100), 0, None)]
    st0 = np.float32(2.0 ** np.floor(np.log2(size / 120)))
    cases.append((None, (128, 120, 136), 0, (np.array([-64 * st0, -60 * st0, -68 * st0], np.float32), st0)))   # on the coordinate planes
    for frac, dims, x_offset, fixed in cases:
        if fixed is None:
            centre = a + (b - a) * rng.uniform(0.3, 0.7, 3)
            st = np.float32(size * frac / max(dims))
            corner = (centre - st * np.array(dims) / 2).astype(np.float32)
        else:
            corner, st = fixed
        _lib.check(L.cc_set_columns_mode(1))
        got = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_columns_mode(0))
        want = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_columns_mode(1))
        assert got.tobytes() == want.tobytes(), "%s %s" % (name, frac)


def test_columns_serve_the_launch(cb, scenes):
    """The column kernels are what runs: the column pass and the brick kernel, the brick centres ahead of them for
    an assembly, the full walk of flagged bricks behind them when a transform row needs the per-column check."""
    from codecad_b200 import _lib
    L = _lib.lib()
    for name, n_launches in (("cfg_planetary", (3, 4)), ("x_gear3d", (2, 3))):
        s = scenes[name]
        prog = s.compiled()
        corner, step = s.grid(64)
        cb.grid_eval(prog, corner, step, (64, 64, 64))
        assert _lib.decode_program(s.words)[0].column_invariant_percent > 0
        n0 = _lib.counters()[0]
        cb.grid_eval(prog, corner, step, (64, 64, 64))
        assert _lib.counters()[0] - n0 in n_launches
        _lib.check(L.cc_set_columns_mode(0))
        n0 = _lib.counters()[0]
        cb.grid_eval(prog, corner, step, (64, 64, 64))
        assert _lib.counters()[0] - n0 < n_launches[0]
        _lib.check(L.cc_set_columns_mode(1))


# ---- the hierarchy sinks through the column kernels (blocks x linear tiles) ----

HIERARCHY_SCENES = ["cfg_airfoil", "cfg_planetary", "x_gear3d", "col_star_x", "col_star_half_turn", "col_crossed_extrusions", "col_assembly",
                    "col_profile_and_sphere", "col_two_levels", "colr_01", "colr_05", "colr_08", "colr_12", "colr_19", "colr_21", "colr_22"]


@pytest.mark.parametrize("name", HIERARCHY_SCENES)
def test_mass_properties_through_columns_is_identical(cb, scenes, name):
    """mass_properties sums are integers: with and without the column kernels they must be the same numbers, and
    the oracle's."""
    import codecad_b200
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from oracle import host
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    a, b = _box(s)
    res = float(max(b - a)) / 150.0
    scene.program_buffer().wait_specialized(ProgramBuffer.SINK_MASS)
    _lib.check(L.cc_set_columns_mode(1))
    n0 = _lib.counters()[0]
    got = codecad_b200.mass_properties(scene, res, 32)
    with_columns = _lib.counters()[0] - n0
    _lib.check(L.cc_set_columns_mode(0))
    n0 = _lib.counters()[0]
    want = codecad_b200.mass_properties(scene, res, 32)
    without = _lib.counters()[0] - n0
    _lib.check(L.cc_set_columns_mode(1))
    if _lib.decode_program(s.words)[0].column_invariant_percent:
        assert with_columns > without                   # the column pass is an extra launch per level chunk: it ran
    assert got.volume == want.volume and tuple(got.centroid) == tuple(want.centroid)
    assert np.array_equal(np.array(got.inertia_tensor), np.array(want.inertia_tensor))
    vol, cen, _ = host.mass_properties(s.words, s.box_a, s.box_b, res, 32)   # (Kahan sums on the host: last-digit differences allowed)
    assert abs(got.volume - vol) <= 1e-12 * abs(vol) and np.allclose(np.array(got.centroid), np.array(cen), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", HIERARCHY_SCENES)
def test_subdivision_through_columns_is_identical(cb, scenes, name):
    import codecad_b200
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    a, b = _box(s)
    res = float(max(b - a)) / 200.0
    scene.program_buffer().wait_specialized(ProgramBuffer.SINK_CLASSIFY)
    out = []
    for mode in (1, 0):
        _lib.check(L.cc_set_columns_mode(mode))
        r = codecad_b200.subdivision(scene, res, grid_size=16)
        out.append(r)
    _lib.check(L.cc_set_columns_mode(1))
    assert tuple(out[0][1]) == tuple(out[1][1])
    assert [tuple(map(tuple, (blk[0], blk[1], blk[3]))) + (blk[2], blk[4]) for blk in out[0][2]] == \
           [tuple(map(tuple, (blk[0], blk[1], blk[3]))) + (blk[2], blk[4]) for blk in out[1][2]]
    assert len(out[0][2]) > 0
