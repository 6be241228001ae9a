"""Part culling: dense float4 grids of an assembly — a tree of sharp unions over
self-contained sub-programs — skip, brick by brick, the parts that provably cannot be the nearest.
Must not change a bit: against the CPU oracle on windows of several scales, and against the same
library with the culling switched off on larger grids."""
import numpy as np
import pytest

import oracle
from scenes import ALL_NAMES

pytestmark = pytest.mark.gpu
PART_SCENES = ["cfg_planetary", "cfg_menger_sponge", "dsdf3d_mirror_3d", "dsdf2d_mirror_2d", "dsdf2d_rotated_pattern_2d",
               "col_assembly", "colr_08", "colr_19", "colr_21"]    # (the last four: random unions of extruded solids, tools/make_fixtures.py --columns)


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    L = _lib.lib()
    old_jit = _lib.check(L.cc_set_jit_mode(2))     # compile at first use and wait: the part kernels are NVRTC kernels
    old_forest = _lib.check(L.cc_set_forest_mode(0))
    yield codecad_b200
    _lib.check(L.cc_set_parts_mode(1))
    _lib.check(L.cc_set_forest_mode(old_forest))
    _lib.check(L.cc_set_jit_mode(old_jit))


def _f4(arr):
    return np.stack([arr["x"], arr["y"], arr["z"], arr["w"]], axis=-1)


def test_which_scenes_have_parts(scenes):
    from codecad_b200 import _lib
    with_parts = sorted(n for n in ALL_NAMES if _lib.decode_program(scenes[n].words)[0].n_parts)
    assert set(PART_SCENES) <= set(with_parts)
    info = _lib.decode_program(scenes["cfg_planetary"].words)[0]
    assert info.n_parts == 7 and info.n_parts_bounded == 7


@pytest.mark.parametrize("name", PART_SCENES)
def test_parts_bit_exact_vs_oracle(cb, scenes, name):
    from codecad_b200 import _lib
    _lib.check(_lib.lib().cc_set_parts_mode(1))
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(3)
    a, b = np.array(s.box_a), np.array(s.box_b)
    if s.dimension == 2:
        a[2], b[2] = -1.0, 1.0
    size = float(max(b - a))
    corner, step = s.grid(48)
    windows = [(corner, step, (48, 40, 33))]
    for frac, dims in ((0.3, (40, 33, 48)), (0.05, (33, 48, 40)), (0.01, (24, 40, 64))):
        for _ in range(2):
            centre = a + (b - a) * rng.uniform(0.2, 0.8, 3)
            st = np.float32(size * frac / 32)
            windows.append(((centre - st * np.array(dims) / 2).astype(np.float32), st, dims))
    for corner, step, dims in windows:
        want = oracle.grid_eval(s.words, corner, step, dims)
        got = _f4(cb.grid_eval(scene, corner, step, dims))
        assert np.array_equal(got, want, equal_nan=True), "%s step %g: %d values differ" % (name, step, int((got != want).sum()))


@pytest.mark.parametrize("name", PART_SCENES)
def test_parts_equal_the_unculled_kernels(cb, scenes, name):
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(17)
    a, b = np.array(s.box_a), np.array(s.box_b)
    if s.dimension == 2:
        a[2], b[2] = -1.0, 1.0
    size = float(max(b - a))
    for frac, dims, x_offset in ((1.05, (128, 128, 128), 0), (0.3, (130, 70, 90), 5), (0.04, (64, 136, 100), 0)):
        centre = a + (b - a) * rng.uniform(0.3, 0.7, 3)
        st = np.float32(size * frac / max(dims))
        corner = (centre - st * np.array(dims) / 2).astype(np.float32)
        _lib.check(L.cc_set_parts_mode(1))
        got = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_parts_mode(0))
        want = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_parts_mode(1))
        assert got.tobytes() == want.tobytes(), "%s frac %g" % (name, frac)


# ---- the same culling on the interpreter tier (cc_parts.cu): what runs before NVRTC has delivered ----

@pytest.fixture()
def interpreter_only(cb):
    from codecad_b200 import _lib
    L = _lib.lib()
    old = _lib.check(L.cc_set_jit_mode(0))
    yield
    _lib.check(L.cc_set_jit_mode(old))


@pytest.mark.parametrize("name", PART_SCENES)
def test_interpreter_parts_bit_exact_vs_oracle(cb, interpreter_only, scenes, name):
    from codecad_b200 import _lib
    _lib.check(_lib.lib().cc_set_parts_mode(1))
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(5)
    a, b = np.array(s.box_a), np.array(s.box_b)
    if s.dimension == 2:
        a[2], b[2] = -1.0, 1.0
    size = float(max(b - a))
    corner, step = s.grid(48)
    windows = [(corner, step, (48, 40, 33))]
    for frac, dims in ((0.3, (40, 33, 48)), (0.05, (33, 48, 40)), (0.01, (24, 40, 64))):
        centre = a + (b - a) * rng.uniform(0.2, 0.8, 3)
        st = np.float32(size * frac / 32)
        windows.append(((centre - st * np.array(dims) / 2).astype(np.float32), st, dims))
    for corner, step, dims in windows:
        want = oracle.grid_eval(s.words, corner, step, dims)
        got = _f4(cb.grid_eval(scene, corner, step, dims))
        assert np.array_equal(got, want, equal_nan=True), "%s step %g: %d values differ" % (name, step, int((got != want).sum()))


@pytest.mark.parametrize("name", PART_SCENES)
def test_interpreter_parts_equal_the_full_walk(cb, interpreter_only, scenes, name):
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    rng = np.random.default_rng(19)
    a, b = np.array(s.box_a), np.array(s.box_b)
    if s.dimension == 2:
        a[2], b[2] = -1.0, 1.0
    size = float(max(b - a))
    for frac, dims, x_offset in ((1.05, (128, 128, 128), 0), (0.3, (130, 70, 90), 5), (0.04, (64, 136, 100), 0)):
        centre = a + (b - a) * rng.uniform(0.3, 0.7, 3)
        st = np.float32(size * frac / max(dims))
        corner = (centre - st * np.array(dims) / 2).astype(np.float32)
        _lib.check(L.cc_set_parts_mode(1))
        n0 = _lib.counters()[0]
        got = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        assert (_lib.counters()[0] - n0) % 2 == 0 and _lib.counters()[0] > n0   # brick centres + the culled walk, per chunk
        _lib.check(L.cc_set_parts_mode(0))
        want = np.array(cb.grid_eval(scene, corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_parts_mode(1))
        assert got.tobytes() == want.tobytes(), "%s frac %g" % (name, frac)


# ---- the hierarchy sinks (blocks x linear tiles): a part mask per tile (cc_tile_centers_body) ----

HIERARCHY_PART_SCENES = ["cfg_planetary", "cfg_menger_sponge", "dsdf3d_mirror_3d", "col_assembly", "colr_08", "colr_21"]


@pytest.mark.parametrize("name", HIERARCHY_PART_SCENES)
@pytest.mark.parametrize("columns", [0, 1])
def test_hierarchy_sinks_with_tile_masks_are_identical(cb, scenes, name, columns):
    """mass_properties (integer sums), subdivision (ordered leaf blocks) and the mesh (triangles of every leaf block)
    with the per-tile part masks against the same calls without them, with and without the column kernels."""
    import codecad_b200
    from codecad_b200 import _lib
    from codecad_b200.cl_util.buffer import ProgramBuffer
    from codecad_b200.rendering import mesh
    L = _lib.lib()
    s = scenes[name]
    scene = s.compiled()
    a, b = np.array(s.box_a), np.array(s.box_b)
    res = float(max(b - a)) / 120.0
    scene.program_buffer().wait_specialized(ProgramBuffer.SINK_MASS | ProgramBuffer.SINK_CLASSIFY | ProgramBuffer.SINK_PYMCUBES)
    old_columns = _lib.check(L.cc_set_columns_mode(columns))
    results, launches = [], []
    try:
        for parts in (1, 0):
            _lib.check(L.cc_set_parts_mode(parts))
            n0 = _lib.counters()[0]
            mp = codecad_b200.mass_properties(scene, res, 32)
            sub = codecad_b200.subdivision(scene, res, grid_size=16)
            vertices, block, _ = mesh.mesh_arrays(scene, 32)
            launches.append(_lib.counters()[0] - n0)
            results.append((mp.volume, tuple(mp.centroid), np.array(mp.inertia_tensor).tobytes(),
                            [tuple(map(tuple, (blk[0], blk[1], blk[3]))) + (blk[2], blk[4]) for blk in sub[2]],
                            np.array(vertices).tobytes(), np.array(block).tobytes()))
    finally:
        _lib.check(L.cc_set_parts_mode(1))
        _lib.check(L.cc_set_columns_mode(old_columns))
    # (a program with a column split has ONE tile unit, whose per-cell body reads the column buffer: with the columns
    # switched off — a debugging switch — its hierarchy sinks run the plain kernels, masks included)
    if columns or not _lib.decode_program(s.words)[0].column_invariant_percent:
        assert launches[0] > launches[1], "the tile-centre passes did not run"
    for k in range(6):
        assert results[0][k] == results[1][k], "%s: result %d differs" % (name, k)
    assert len(results[0][3]) > 0 and len(results[0][4]) > 0
