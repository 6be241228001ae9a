"""Union forests (csrc/cc_forest.cu): dense grids of a program that is one tree of unions over fused
primitives are evaluated tile by tile with the far primitives culled — and must not differ from the
full evaluation in a single bit: against the CPU oracle on windows of every scale (a tile that
spans the whole scene, a tile much smaller than a primitive), and against the library's own
unculled kernels on larger grids."""
import numpy as np
import pytest

import oracle
from scenes import FOREST_NAMES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    old = _lib.check(_lib.lib().cc_set_jit_mode(0))   # unculled reference inside the library = the interpreter
    yield codecad_b200
    _lib.check(_lib.lib().cc_set_forest_mode(1))
    _lib.check(_lib.lib().cc_set_jit_mode(old))


def _f4(arr):
    return np.stack([arr["x"], arr["y"], arr["z"], arr["w"]], axis=-1)


def _windows(scene, rng):
    """(corner, step, dims): the whole box coarsely, then windows at finer and finer steps placed on
    primitives (centres of the box and random spots), with ragged dims so that edge tiles occur."""
    a, b = np.array(scene.box_a), np.array(scene.box_b)
    size = float(max(b - a))
    corner, step = scene.grid(32)
    out = [(corner, step, (20, 24, 32)), (corner, step, (33, 17, 40))]
    for frac, dims in ((0.25, (40, 33, 48)), (0.05, (48, 40, 35)), (0.01, (37, 48, 64)), (0.002, (32, 32, 48))):
        for _ in range(2):
            centre = a + (b - a) * rng.uniform(0.2, 0.8, 3)
            st = np.float32(size * frac / 32)
            out.append(((centre - st * np.array(dims) / 2).astype(np.float32), st, dims))
    return out


@pytest.mark.parametrize("name", FOREST_NAMES)
def test_forest_grid_eval_bit_exact_vs_oracle(cb, scenes, name):
    from codecad_b200 import _lib
    s = scenes[name]
    assert s.compiled().program_buffer().info.n_forest_leaves > 0, "not recognised as a union forest"
    _lib.check(_lib.lib().cc_set_forest_mode(1))
    rng = np.random.default_rng(5)
    windows = _windows(s, rng)
    if name == "cfg_synthetic500":
        windows = windows[:1] + windows[2::2]       # the oracle walks 500 boxes per point
    for corner, step, dims in windows:
        want = oracle.grid_eval(s.words, corner, step, dims)
        got = _f4(cb.grid_eval(s.compiled(), corner, step, dims))
        same = np.array_equal(got, want, equal_nan=True)
        assert same, "%s window step %g: %d of %d values differ" % (name, step, int((got != want).sum()), got.size)


@pytest.mark.parametrize("name", FOREST_NAMES)
def test_forest_equals_the_unculled_kernels(cb, scenes, name):
    """Larger grids, GPU against GPU: the same call with the culling kernel switched off."""
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes[name]
    rng = np.random.default_rng(11)
    a, b = np.array(s.box_a), np.array(s.box_b)
    size = float(max(b - a))
    for frac, dims, x_offset in ((1.05, (96, 96, 96), 0), (0.2, (130, 70, 90), 7), (0.03, (64, 128, 96), 0)):
        centre = a + (b - a) * rng.uniform(0.3, 0.7, 3)
        st = np.float32(size * frac / max(dims))
        corner = (centre - st * np.array(dims) / 2).astype(np.float32)
        _lib.check(L.cc_set_forest_mode(1))
        got = np.array(cb.grid_eval(s.compiled(), corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_forest_mode(0))
        want = np.array(cb.grid_eval(s.compiled(), corner, st, dims, x_offset=x_offset))
        _lib.check(L.cc_set_forest_mode(1))
        assert got.tobytes() == want.tobytes(), "%s frac %g" % (name, frac)


def test_forest_counts_as_one_launch_and_needs_no_compilation(cb, scenes):
    from codecad_b200 import _lib
    L = _lib.lib()
    s = scenes["cfg_synthetic500"]
    corner, step = s.grid(64)
    L.cc_reset_counters()
    cb.grid_eval(s.compiled(), corner, step, (64, 64, 64))
    launches, points = _lib.counters()
    assert launches == 1 and points == 64 ** 3
