"""Host-side logic of the drop-in modules (no GPU): block planner, pipelines, slabs, the
mass-property post-processing."""
import itertools
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys

import numpy as np
import pytest

from codecad_b200 import BoundingBox, Vector, calculate_block_sizes
from codecad_b200 import mass_properties as mp_mod_fn  # noqa: F401 (function re-export exists)
from codecad_b200.cl_util import interleave, interleave2
from codecad_b200.grid_eval import slab_range
from oracle import host

import importlib
mp_mod = importlib.import_module("codecad_b200.mass_properties")


# ---- calculate_block_sizes: reference tests/test_subdivision.py:44-107 (80 cases) ------------
@pytest.mark.parametrize("box_size", [Vector(10, 20, 30), Vector(16, 16, 16)])
@pytest.mark.parametrize("dimension", [2, 3])
@pytest.mark.parametrize("resolution", [1, 0.1])
@pytest.mark.parametrize("grid_size, multiplier", [(2, 1), (2, 2), (21, 1), (256, 1), (256, 256)])
@pytest.mark.parametrize("overlap", [True, False])
def test_block_sizes(box_size, dimension, resolution, grid_size, overlap, multiplier):
    box = BoundingBox(-box_size / 2, box_size / 2)
    bs = calculate_block_sizes(box, dimension, resolution, grid_size, overlap, multiplier)

    assert bs[-1][0] == 1, "Final block size must have block size 1"
    for i, (level_resolution, level_size) in enumerate(bs):
        if i > 0:
            for j in range(dimension):
                assert level_size[j] == grid_size
        if dimension == 2:
            assert level_size[2] == 1
        assert level_size[0] > 1 or level_size[1] > 1 or level_size[2] > 1
        for j in range(dimension):
            assert level_size[j] % multiplier == 0

    real_block_size = bs[0][0] * resolution
    if overlap and len(bs) == 1:
        for i in range(dimension):
            assert bs[0][1][i] >= box_size[i] / real_block_size + 1
            assert multiplier > 1 or box_size[i] / real_block_size + 1 > (bs[0][1][i] - 1)
    else:
        for i in range(dimension):
            assert bs[0][1][i] >= box_size[i] / (bs[0][0] * resolution)
            assert multiplier > 1 or box_size[i] / real_block_size > (bs[0][1][i] - 1)

    for level_number, ((lr, ls), (pr, ps)) in enumerate(zip(bs[:-1], bs[1:])):
        for i in range(dimension):
            if overlap and level_number == len(bs) - 2:
                assert lr == pr * (ps[i] - 1)
            else:
                assert lr == pr * ps[i]

    # and it is the same plan as the oracle's restatement of subdivision.py:116-166
    want = host.calculate_block_sizes(tuple(box.a), tuple(box.b), dimension, resolution, grid_size, overlap, multiplier)
    assert [(c, tuple(d)) for c, d in bs] == [(c, tuple(d)) for c, d in want]


def test_block_sizes_rejects_bad_multiplier():
    with pytest.raises(ValueError):
        calculate_block_sizes(BoundingBox(Vector(0, 0, 0), Vector(1, 1, 1)), 3, 0.1, 10, False, 3)


def test_block_sizes_survey_values():
    """SURVEY.md 8(a6) probed plans for csg_example at resolution 100/512"""
    box = BoundingBox(Vector(-50, -50, -50), Vector(50, 50, 50)).expanded_additive(100 / 512 / 2)
    bs = calculate_block_sizes(box, 3, 100 / 512, 128, True)
    assert [(c, tuple(d)) for c, d in bs] == [(127, (5, 5, 5)), (1, (128, 128, 128))]
    bs = calculate_block_sizes(box, 3, 100 / 512, 16, True)
    assert [(c, tuple(d)) for c, d in bs] == [(240, (3, 3, 3)), (15, (16, 16, 16)), (1, (16, 16, 16))]


def test_argument_contracts_match_the_reference():
    """subdivision.py:204-208 and mass_properties.py:38-41 raise AssertionError before any
    device work"""
    import codecad_b200
    from scenes import load_scenes
    s = load_scenes()["mp_unit_box"].compiled()
    s._buffer = object()  # must not be reached: no GPU here
    for kw in ({"resolution": 0}, {"resolution": 1, "grid_size": 1}, {"resolution": 1, "grid_size": 257}):
        with pytest.raises(AssertionError):
            codecad_b200.subdivision(s, **kw)
    for kw in ({"resolution": -1}, {"resolution": 1, "grid_size": 1}, {"resolution": 1, "grid_size": 85}):
        with pytest.raises(AssertionError):
            codecad_b200.mass_properties(s, **kw)
    flat = load_scenes()["dsdf2d_circle"].compiled()
    with pytest.raises(AssertionError):
        codecad_b200.mass_properties(flat, 0.1)


# ---- interleave2: reference tests/test_clutil.py:188-248 -------------------------------------------
def test_interleave2_keeps_two_jobs_in_flight():
    job_log = []

    class MockEvent:
        def __init__(self, job):
            self.job = job

        def wait(self):
            job_log.append((self.job, "wait"))

    def job_func(job):
        job_log.append((job, 1))
        yield MockEvent(job)
        job_log.append((job, 2))
        if job < 10:
            return [2 * job, 2 * job + 1]
        return None

    interleave2(job_func, [2, 3])

    checked = set()
    working = set()
    tics_working = 0
    max_working = 0
    for job, step in job_log:
        if working:
            tics_working += 1
        max_working = max(max_working, len(working))
        if step == 1:
            working.add(job)
        elif step == "wait":
            working.remove(job)
        elif step == 2:
            assert (job, "wait") in checked
        if job > 3:
            assert (job // 2, 2) in checked, "Parent task must be finished before running a child"
        checked.add((job, step))
    assert (18, 2) in checked and (19, 2) in checked
    assert max_working == 2
    assert tics_working >= len(checked) - 4


def test_interleave_helpers_alternate():
    log = []

    class Helper:
        def __init__(self, name):
            self.name = name

        def enqueue(self, job):
            log.append((self.name, "enqueue", job))
            self.job = job
            return job

        def process_result(self, event):
            log.append((self.name, "result", event))
            return [(event * 2,), (event * 2 + 1,)] if event < 4 else []

    interleave([(1,)], Helper("a"), Helper("b"))   # jobs are argument tuples (cl_buffer.py:182)
    done = sorted(j for _, what, j in log if what == "result")
    assert done == [1, 2, 3, 4, 5, 6, 7]
    # a job's result is processed by the helper that enqueued it
    for name, what, job in log:
        if what == "result":
            assert (name, "enqueue", job) in log


# ---- slabs ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 8), (513, 4), (7, 8), (2048, 8)])
def test_slab_range_partitions_the_axis(n, world):
    edges = [slab_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


# ---- mass-property post-processing (mass_properties.py:179-229) ---------------------------------------
def test_finish_matches_oracle_and_analytic_box():
    # integrals of the box [1,3] x [0,2] x [-1,5]
    a, b = np.array([1.0, 0.0, -1.0]), np.array([3.0, 2.0, 5.0])
    size = b - a
    vol = size.prod()
    first = vol * (a + b) / 2
    second = vol * (a * a + a * b + b * b) / 3
    c = (a + b) / 2
    ints = [vol, first[0], first[1], first[2], second[0], second[1], second[2],
            vol * c[0] * c[1], vol * c[0] * c[2], vol * c[1] * c[2]]
    got = mp_mod.finish(ints)
    want = host.finish_mass_properties(ints)
    assert got.volume == want[0] and tuple(got.centroid) == tuple(want[1])
    assert np.array_equal(got.inertia_tensor, want[2])
    assert got.volume == pytest.approx(24.0)
    assert tuple(got.centroid) == pytest.approx((2.0, 1.0, 2.0))
    ex = vol / 12 * np.array([size[1] ** 2 + size[2] ** 2, size[0] ** 2 + size[2] ** 2, size[0] ** 2 + size[1] ** 2])
    assert np.allclose(np.diag(got.inertia_tensor), ex)
    assert np.allclose(got.inertia_tensor - np.diag(np.diag(got.inertia_tensor)), 0, atol=1e-9)


def test_finish_zero_volume():
    r = mp_mod.finish([0.0] * 10)
    assert r.volume == 0 and tuple(r.centroid) == (0, 0, 0) and np.array_equal(r.inertia_tensor, np.zeros((3, 3)))


def test_vector_matches_reference_semantics():
    v = Vector(1, 2)
    assert v.z == 0 and v == (1, 2, 0) and hash(v) == hash((1, 2, 0))
    assert (Vector(1, 2, 3) * 2 + Vector(1, 1, 1)) == (3, 5, 7)
    f4 = Vector(0.1, 0.2, 0.3).as_float4()
    assert f4.dtype.names == ("x", "y", "z", "w") and f4["x"] == np.float32(0.1) and f4["w"] == 0


def test_leaf_blocks_sequence_is_the_reference_list():
    """subdivision()'s block list is built on demand; it must behave like the reference's list of
    (dims, corner Vector, step, int_corner Vector, int_step) tuples (subdivision.py:97-111)."""
    import numpy as np
    from codecad_b200.geometry import Vector
    from codecad_b200.subdivision import LeafBlocks, block_corners
    ints = np.array([[0, 15, 30], [45, 0, 15], [15, 15, 15]], dtype=np.int64)
    pos = ints * 0.25 + np.array([-1.0, -2.0, -3.0])
    dims = Vector(16, 16, 16)
    blocks = LeafBlocks(dims, pos, 0.25, ints, 1)
    want = [(dims, Vector(*p), 0.25, Vector(*i), 1) for p, i in zip(pos.tolist(), ints.tolist())]
    assert len(blocks) == 3 and blocks[1] == want[1] and blocks[-1] == want[-1]     # before materialising
    assert isinstance(blocks[0][1], Vector) and isinstance(blocks[0][3].x, int)
    assert list(blocks) == want and blocks == want and blocks[0:2] == want[0:2]
    acc = []
    acc += blocks
    assert acc == want and sorted(tuple(b[3]) for b in blocks) == sorted(tuple(i) for i in ints.tolist())
    for size, corner, step, int_corner, int_step in blocks:
        assert size == dims and step == 0.25 and int_step == 1
    with __import__("pytest").raises(IndexError):
        blocks[3]
    assert np.array_equal(block_corners(blocks), pos) and np.array_equal(block_corners(want), pos)
    assert len(LeafBlocks(dims, pos[:0], 0.25, ints[:0], 1)) == 0 and not list(LeafBlocks(dims, pos[:0], 0.25, ints[:0], 1))


# ---- exact accumulator of the mass integrals and the canonical leaf order (host-only entry points) ----

def _to_limbs(value, exponent):
    """trunc(value / 2^exponent) cut into four signed 32-bit limbs (what cc_mass_integrals_kernel adds up)."""
    from fractions import Fraction
    t = int(Fraction(value) / Fraction(2) ** exponent)        # int() truncates toward zero
    sign, mag = (-1 if t < 0 else 1), abs(t)
    return [sign * ((mag >> (32 * k)) & 0xFFFFFFFF) for k in range(4)]


def test_limb_sums_convert_exactly_and_order_independently():
    import random
    from fractions import Fraction
    import importlib
    mpm = importlib.import_module("codecad_b200.mass_properties")
    rng = random.Random(7)
    exps = np.array([-60, -55, -55, -55, -50, -50, -50, -50, -50, -50], dtype=np.int32)
    values = [[rng.uniform(-1, 1) * 10 ** rng.uniform(-6, 6) for _ in range(10)] for _ in range(300)]
    def total(rows):
        limbs = np.zeros(40, dtype=np.int64)
        for row in rows:
            for i, v in enumerate(row):
                limbs[4 * i:4 * i + 4] += np.array(_to_limbs(v, int(exps[i])), dtype=np.int64)
        return limbs
    a = total(values)
    rng.shuffle(values)
    b = total(values[:100]) + total(values[100:])            # another order, another partition
    assert np.array_equal(a, b)
    got = mpm.limbs_to_integrals(a, exps)
    for i in range(10):
        exact = sum(Fraction(int(Fraction(row[i]) / Fraction(2) ** int(exps[i]))) for row in values) * Fraction(2) ** int(exps[i])
        assert got[i] == float(exact)                          # one correctly rounded conversion
        assert abs(got[i] - sum(row[i] for row in values)) <= 1e-9 * sum(abs(row[i]) for row in values)


def test_sort_leaf_corners_restores_the_single_device_order():
    import random
    from codecad_b200.geometry import Vector
    from codecad_b200.subdivision import sort_leaf_corners
    rng = random.Random(3)
    for grid, overlap in ((16, True), (8, False), (5, True)):
        leaf_factor = grid - 1 if overlap else grid
        plan = [(leaf_factor * grid, Vector(3, 2, 3)), (leaf_factor, Vector.splat(grid)), (1, Vector.splat(grid))]
        # breadth-first production: level-0 hits in INDEX3 order, then per parent the level-1 hits in INDEX3 order
        def hits(dims, frac):
            cells = [(x, y, z) for x in range(dims[0]) for y in range(dims[1]) for z in range(dims[2])]
            return [c for c in cells if rng.random() < frac]
        corners = []
        for c0 in hits((3, 2, 3), 0.6):
            for c1 in hits((grid,) * 3, 0.05):
                corners.append([c0[k] * plan[0][0] + c1[k] * plan[1][0] for k in range(3)])
        want = np.array(corners, dtype=np.int64)
        order = list(range(len(want)))
        rng.shuffle(order)
        shuffled = want[order]
        # dealing to ranks permutes the list; the sort must undo any permutation
        assert np.array_equal(sort_leaf_corners(shuffled, plan), want)
        assert np.array_equal(sort_leaf_corners(want[::-1], plan), want)


def test_executed_flops_file_matches_the_instrumented_oracle(scenes):
    """profiles/executed_flops.json (what bench.py reports as flop_per_point_executed) is the oracle's
    executed-branch count; re-measured here on a coarser subsample, and bracketed by the loader's
    static minimum / maximum."""
    import json
    import oracle
    from codecad_b200 import _lib
    data = json.load(open(os.path.join(ROOT, "profiles", "executed_flops.json")))["scenes"]
    for name in ("cfg_csg_example", "cfg_menger_sponge", "cfg_airfoil", "cfg_planetary"):
        s = scenes[name]
        n = data[name]["grid"]
        corner, step = s.grid(n)
        mean, pts = oracle.executed_flops(s.words, corner, step, (n, n, n), stride=n // 32)
        info, _ = _lib.decode_program(s.words)
        assert info.flops_min <= data[name]["flop_per_point_executed"] <= info.flops_max
        assert mean == pytest.approx(data[name]["flop_per_point_executed"], rel=0.02)


def test_leaf_block_list_built_in_c_equals_the_python_form():
    """csrc/cc_pylist.c builds subdivision()'s list of block tuples; same objects as the Python form."""
    import importlib
    sub = importlib.import_module("codecad_b200.subdivision")
    if sub._pylist is None:
        pytest.skip("codecad_b200/_cc_pylist.so not built")
    rng = np.random.default_rng(2)
    ic = rng.integers(-5000, 5000, (777, 3)).astype(np.int64)
    c = ic * 0.195 + (-50.1)
    dims = Vector(16, 16, 16)
    fast = list(sub.LeafBlocks(dims, c, 0.195, ic, 15))
    saved, sub._pylist = sub._pylist, None
    try:
        slow = list(sub.LeafBlocks(dims, c, 0.195, ic, 15))
    finally:
        sub._pylist = saved
    assert fast == slow and len(fast) == 777
    b = fast[3]
    assert type(b) is tuple and type(b[1]) is Vector and type(b[3]) is Vector
    assert type(b[1].x) is float and type(b[3].z) is int and b[0] is dims and b[4] == 15
    assert sub.LeafBlocks(dims, c[:0], 0.195, ic[:0], 15)._materialise() == []


def test_cut_equal_work():
    """grid_eval.cut_equal_work: the pure part of balanced_slabs."""
    import importlib
    ge = importlib.import_module("codecad_b200.grid_eval")
    assert ge.cut_equal_work([3.0] * 16, 4) is None                       # nothing to balance
    assert ge.cut_equal_work([1.0, 2.0], 3) is None                       # fewer layers than ranks
    cost = np.array([1, 1, 1, 1, 4, 4, 4, 4, 1, 1, 1, 1], float)          # a heavy middle
    cuts = ge.cut_equal_work(cost, 3)
    assert cuts[0] == 0 and cuts[-1] == len(cost) and all(b > a for a, b in zip(cuts, cuts[1:]))
    sums = [cost[a:b].sum() for a, b in zip(cuts, cuts[1:])]
    assert max(sums) <= 1.5 * min(sums)
    assert cuts[1] > len(cost) // 3 and cuts[2] < 2 * len(cost) // 3      # the rim runs are longer than the middle one
    rng = np.random.default_rng(5)
    for _ in range(200):
        n, w = int(rng.integers(2, 200)), int(rng.integers(1, 17))
        c = rng.uniform(0.0, 10.0, n) ** 3
        cuts = ge.cut_equal_work(c, w)
        if n < w:
            assert cuts is None
            continue
        assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == w + 1 and all(b > a for a, b in zip(cuts, cuts[1:]))


def test_row_with_residue_coefficient_is_monotone_along_the_column():
    """The run-time check of the column kernels: a transform row is a chain of correctly rounded FMAs,
    hence monotone in the coordinate along the column; equal bits at both ends of a column therefore mean equal bits at
    every cell between.  Checked here on random rows with a residue coefficient (fp32 FMA emulated exactly in float64:
    the product of two floats is exact there, and rounding the sum twice keeps it monotone)."""
    rng = np.random.default_rng(77)

    def fma(a, b, c):
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))

    constant = varying = 0
    for _ in range(400):
        mx, my = np.float32(rng.uniform(-1, 1)), np.float32(rng.uniform(-1, 1))
        mz = np.float32(rng.choice([-1, 1]) * 10.0 ** rng.uniform(-18, -9))
        o = np.float32(rng.choice([0.0, 1e-14, -8.4e-15, rng.uniform(-50, 50)]))
        step, cz = np.float32(10.0 ** rng.uniform(-3, 0)), np.float32(rng.uniform(-60, 60))
        x = np.float32(rng.choice([0.0, rng.uniform(-60, 60)]))
        y = np.float32(rng.choice([0.0, rng.uniform(-60, 60), 10.0 ** rng.uniform(-14, -6)]))
        zs = [fma(step, np.float32(i), cz) for i in range(64)]
        row = np.array([fma(mx, x, fma(my, y, fma(mz, z, o))) for z in zs], np.float32)
        d = np.diff(row.astype(np.float64))
        assert np.all(d >= 0) or np.all(d <= 0), "not monotone"
        if row[0].tobytes() == row[-1].tobytes():
            assert all(v.tobytes() == row[0].tobytes() for v in row)
            constant += 1
        else:
            varying += 1
    assert constant > 100 and varying > 5          # both outcomes occur: the check is neither vacuous nor always failing
