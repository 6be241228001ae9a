"""Sharding of the hierarchy paths, on ONE GPU: the share of every rank is
computed by the same device in turn, and the shares must add up to EXACTLY the unsharded result —
bit-identical mass integrals (exact accumulator, cc_mass_properties_exact) and, after
cc_sort_leaf_corners, the identical leaf-block array.  tests/test_gpu_multi.py repeats this with
real ranks / devices where two GPUs are present."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MASS = [("cfg_airfoil", 1.0, 64), ("cfg_airfoil", 0.5, 64), ("cfg_csg_example", 1.0, 16), ("cfg_csg_example", 0.5, 8),
        ("cfg_menger_sponge", 0.8, 32), ("mp_drunk_box", 0.05, 8), ("mp_sphere", 0.02, 4)]
SUBDIV = [("cfg_csg_example", 100 / 512, 16, True), ("cfg_csg_example", 100 / 128, 8, False), ("sub_box10", 1.0, 4, True),
          ("cfg_menger_sponge", 0.75, 16, True), ("dsdf2d_gear", 0.05, 8, True), ("cfg_csg_example", 100 / 512, 128, True)]


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    return codecad_b200


def _plan(scene, resolution, grid, overlap, expand):
    from codecad_b200.geometry import BoundingBox, as_vector
    from codecad_b200.subdivision import calculate_block_sizes
    box = BoundingBox(as_vector(scene.box_a), as_vector(scene.box_b))
    if expand:
        box = box.expanded_additive(resolution / 2)
        if scene.dimension == 2:
            box = box.flattened()
    return box, calculate_block_sizes(box, scene.dimension, resolution, grid, overlap)


@pytest.mark.parametrize("name,resolution,grid", MASS)
@pytest.mark.parametrize("world", [2, 3, 8])
def test_mass_shares_add_up_exactly(cb, scenes, name, resolution, grid, world):
    mpm = importlib.import_module("codecad_b200.mass_properties")
    s = scenes[name]
    prog = s.compiled().program_buffer()
    box, plan = _plan(s, resolution, grid, False, False)
    whole, exps, st = mpm.mass_limbs(prog, box.a, resolution, plan)
    total = np.zeros(40, dtype=np.int64)
    cells = []
    for rank in range(world):
        limbs, e, st_r = mpm.mass_limbs(prog, box.a, resolution, plan, rank, world)
        assert np.array_equal(e, exps)
        total += limbs
        cells.append(st_r["cells"])
    assert np.array_equal(total, whole), "the shares' accumulators must add up to the unsharded one, limb by limb"
    assert np.array_equal(mpm.limbs_to_integrals(total, exps), mpm.limbs_to_integrals(whole, exps))
    # the deepest level dominates and is dealt round-robin: no share does much more than its part
    if len(plan) >= 2 and st["cells"] > 50 * world * grid ** 3:
        assert max(cells) <= 1.35 * (sum(cells) / world)


@pytest.mark.parametrize("name,resolution,grid,overlap", SUBDIV)
@pytest.mark.parametrize("world", [2, 5])
def test_subdivision_shares_merge_into_the_single_device_order(cb, scenes, name, resolution, grid, overlap, world):
    sub = importlib.import_module("codecad_b200.subdivision")
    s = scenes[name]
    prog = s.compiled().program_buffer()
    box, plan = _plan(s, resolution, grid, overlap, True)
    if len(plan) < 2:
        pytest.skip("single-level plan")
    whole = sub.subdivide_int_corners(prog, box.a, resolution, plan, s.dimension)
    parts = [sub.subdivide_int_corners(prog, box.a, resolution, plan, s.dimension, r, world) for r in range(world)]
    assert sum(len(p) for p in parts) == len(whole)
    merged = sub.sort_leaf_corners(np.concatenate(parts), plan)
    assert np.array_equal(merged, whole)
    assert np.array_equal(sub.sort_leaf_corners(whole, plan), whole), "one device already lists blocks in that order"


def test_exact_accumulator_matches_a_float64_sum(cb, scenes):
    """The exact accumulator changes the LAST bits of the reference's Kahan sum at most."""
    from oracle import host
    s = scenes["cfg_airfoil"]
    vol, cen, inertia = host.mass_properties(s.words, s.box_a, s.box_b, 1.0, 64)
    got = cb.mass_properties(s.compiled(), 1.0, 64)
    assert abs(got.volume - vol) <= 4e-16 * vol
    assert max(abs(g - w) for g, w in zip(got.centroid, cen)) <= 1e-13 * max(map(abs, cen))


def test_balanced_slabs_cover_the_grid_and_do_not_change_results(scenes):
    """grid_eval.balanced_slabs: cuts on multiples of eight planes, every plane exactly once, more planes where the
    assembly's box is empty; the slabs' results are the unsharded grid."""
    import importlib
    import codecad_b200
    ge = importlib.import_module("codecad_b200.grid_eval")
    s = scenes["cfg_planetary"]
    scene = s.compiled()
    n = 256
    corner, step = s.grid(n)
    for world in (2, 3, 8):
        slabs = ge.balanced_slabs(scene, corner, step, (n, n, n), world)
        assert len(slabs) == world and slabs[0][0] == 0 and slabs[-1][1] == n
        assert all(a1 == b0 for (_, a1), (b0, _) in zip(slabs, slabs[1:]))
        assert all(x1 > x0 and x0 % 8 == 0 for x0, x1 in slabs)
    slabs = ge.balanced_slabs(scene, corner, step, (n, n, n), 8)
    widths = [x1 - x0 for x0, x1 in slabs]
    assert widths[0] > widths[3] and widths[-1] > widths[4]          # the rim slabs are wider than the middle ones
    whole = np.array(codecad_b200.grid_eval(scene, corner, step, (n, n, n)))
    for x0, x1 in slabs:
        part = np.array(codecad_b200.grid_eval(scene, corner, step, (x1 - x0, n, n), x_offset=x0))
        assert part.tobytes() == whole[x0:x1].tobytes()
    # a program without parts: equal slabs
    c = scenes["cfg_csg_example"]
    cc, cs = c.grid(128)
    assert ge.balanced_slabs(c.compiled(), cc, cs, (128, 128, 128), 4) == [ge.slab_range(128, r, 4) for r in range(4)]
