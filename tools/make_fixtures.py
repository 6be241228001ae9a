#!/usr/bin/env python3
"""Generate tests/golden/scenes.npz from the reference's own Python front end.

Runs ONLY in the build container (needs /root/reference, which does not exist on
the GPU box).  It imports the reference's shape API and node compiler
(codecad/nodes/program.py:74-76 `make_program`) with the import-time stubs in
tools/refstub/ standing in for pyopencl / py-flags, and stores for every scene:

  <name>.words      float32 program words exactly as the reference emits them
  <name>.meta       float64 [dimension, bbox.a.xyz, bbox.b.xyz, feature_size]

`random.seed(0)` is set immediately before every make_program() because the
reference scheduler shuffles with the unseeded global RNG
(codecad/nodes/scheduler.py:165-178, SURVEY.md §7 hard part 4).

Usage:  python tools/make_fixtures.py            (writes tests/golden/scenes.npz)
"""
import math
import os
import random
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("CODECAD_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(HERE, "refstub"))
sys.path.insert(1, REF)
sys.path.insert(2, os.path.join(REF, "examples"))
sys.path.insert(3, os.path.join(REF, "tests"))
warnings.simplefilter("ignore")

import numpy  # noqa: E402

import codecad  # noqa: E402
import codecad.shapes as s  # noqa: E402
from codecad.shapes import simple2d, polygons2d  # noqa: E402

# SURVEY.md §7 hard part 8: examples/airfoil.py is broken at reference HEAD
# (shapes/airfoils.py:9 names simple2d.Polygon2D).  Harness-side alias only.
simple2d.Polygon2D = polygons2d.Polygon2D


def synthetic_deep_csg(n=500, seed=1234):
    """BASELINE.json config 5 / SURVEY.md §8(d) C5: n rounded boxes, rotated and
    translated, joined by a balanced binary tree of smooth unions (r=0.5)."""
    rng = random.Random(seed)
    parts = []
    for _ in range(n):
        b = s.box(rng.uniform(2, 8), rng.uniform(2, 8), rng.uniform(2, 8))
        b = b.offset(rng.uniform(0.1, 0.5))
        axis = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1))
        b = b.rotated(axis, rng.uniform(0, 360))
        b = b.translated(rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-50, 50))
        parts.append(b)
    while len(parts) > 1:
        nxt = []
        for i in range(0, len(parts) - 1, 2):
            nxt.append(s.union([parts[i], parts[i + 1]], r=0.5))
        if len(parts) % 2:
            nxt.append(parts[-1])
        parts = nxt
    return parts[0]


def forest_scenes():
    """Union forests for csrc/cc_forest.cu beyond config C5: cylinders as well as boxes, axis-aligned
    and scaled placements, sharp and rounded unions of several radii in one tree, a left-deep chain
    (an n-ary union), and a dense cluster in which most primitives overlap."""
    out = {}
    rng = random.Random(99)

    def prim(spread, allow_scale=True):
        kind = rng.random()
        if kind < 0.5:
            p = s.box(rng.uniform(1, 6), rng.uniform(1, 6), rng.uniform(1, 6))
        else:
            p = s.cylinder(h=rng.uniform(1, 6), r=rng.uniform(0.5, 3))
        if rng.random() < 0.6:
            p = p.offset(rng.uniform(0.05, 0.6))
        if allow_scale and rng.random() < 0.25:
            p = p.scaled(rng.uniform(0.5, 2.0))
        if rng.random() < 0.7:
            p = p.rotated((rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)), rng.uniform(0, 360))
        elif rng.random() < 0.5:
            p = p.rotated((0, 0, 1), 90)
        return p.translated(rng.uniform(-spread, spread), rng.uniform(-spread, spread), rng.uniform(-spread, spread))

    def tree(parts, radii):
        parts = list(parts)
        while len(parts) > 1:
            i = rng.randrange(len(parts) - 1)
            parts[i:i + 2] = [s.union([parts[i], parts[i + 1]], r=rng.choice(radii))]
        return parts[0]

    out["forest_mixed40"] = tree([prim(20) for _ in range(40)], [0.5, 0.2, -1, 1.0])
    out["forest_sharp24"] = tree([prim(12) for _ in range(24)], [-1])
    out["forest_dense64"] = tree([prim(6) for _ in range(64)], [0.5, 0.3])
    out["forest_chain12"] = s.union([prim(10) for _ in range(12)], r=0.3)
    out["forest_big_r16"] = tree([prim(15, False) for _ in range(16)], [3.0, 0.1])
    return out


def column_scenes():
    """Extrusions for the column kernels beyond the configs: profiles under mirror / symmetry /
    offset / shell, nested transformations, extrusion axes along x, y, z and oblique, half turns (whose
    quaternions leave rounding residue in the matrix), profiles combined with solids that do see every axis,
    an intersection of two extrusions along different axes, extrusions under repetition and revolution."""
    out = {}
    star = s.polygon2d([(0, 3), (0.8, 0.9), (3, 0.7), (1.2, -0.6), (1.9, -2.8), (0, -1.4), (-1.9, -2.8), (-1.2, -0.6), (-3, 0.7),
                        (-0.8, 0.9)])
    gear = s.gears.InvoluteGear(11, 0.8)
    hexagon = s.regular_polygon2d(6, r=2.5)
    out["col_star_z"] = star.extruded(3)
    out["col_star_y"] = star.extruded(3).rotated((1, 0, 0), 90)
    out["col_star_x"] = star.extruded(3).rotated((0, 1, 0), 90).translated(0.5, -0.25, 1)
    out["col_star_half_turn"] = star.extruded(2).rotated((1, 0, 0), 180).translated(0, 0.5, 0)
    out["col_star_oblique"] = star.extruded(3).rotated((1, 2, 3), 25)
    out["col_gear_rot_z"] = gear.extruded(2).rotated((0, 0, 1), 33).translated(1, 2, 0.5)
    out["col_hexagon_shell"] = hexagon.shell(0.3).extruded(4).translated(0, 0, 1)
    out["col_offset_mirror"] = (star.offset(0.2).translated(4, 0) + star.mirrored_x().translated(-4, 0)).extruded(2.5)
    out["col_symmetrical"] = (s.rectangle(2, 1).translated(2, 0.5) + s.circle(d=1.5).translated(3, 2)).symmetrical_x().extruded(1.5)
    out["col_profile_and_sphere"] = star.extruded(2) + s.sphere(d=3).translated(0, 0, 2) - s.cylinder(h=10, d=0.8)
    out["col_crossed_extrusions"] = hexagon.extruded(8) & gear.scaled(1.5).extruded(8).rotated((1, 0, 0), 90)
    out["col_two_levels"] = (star.scaled(0.5).rotated(20).translated(1, 1) & s.circle(d=2.6).translated(1, 1)).extruded(2).rotated((0, 0, 1), 45) \
        .translated(0, 0, -1).scaled(1.5)
    out["col_revolved_and_extruded"] = s.rectangle(1, 2).translated(3, 0).revolved() + star.scaled(0.6).extruded(5)
    out["col_repeated"] = codecad.shapes.unsafe.Repetition(s.circle(d=1.2).extruded(1), (2.5, 2.5, None)) & s.box(9, 9, 3)
    out["col_bolt_circle"] = hexagon.scaled(1.6).extruded(1) - codecad.shapes.unsafe.CircularRepetition(s.circle(d=0.9).translated(2.6, 0).extruded(3), 6)
    out["col_hole_grid"] = s.rectangle(9, 7).extruded(1.2) - codecad.shapes.unsafe.Repetition(s.circle(d=0.8).extruded(4), (1.5, 1.5, None)) \
        + star.scaled(0.4).extruded(3).translated(0, 0, 1)
    out["col_assembly"] = s.union([gear.extruded(1).translated(-4, 0, 0), hexagon.extruded(2).translated(4, 0, 0.5),
                                   star.scaled(0.7).extruded(1.5).rotated((0, 0, 1), 10).translated(0, 5, 0), s.sphere(d=2).translated(0, -5, 0)])
    out.update(random_column_scenes())
    return out


def random_column_scenes(n=24, seed=4242):
    """Random compositions for the column analysis: 2-D shapes under random stacks of mirror / symmetry / offset / shell /
    rotation / translation / scaling and CSG, extruded, turned by multiples of 90 degrees about random axes (half and
    quarter turns leave rounding residue in the matrices) or by arbitrary angles about z, then combined with other such
    solids and with spheres / boxes in general position."""
    rng = random.Random(seed)

    def shape2d(depth=0):
        k = rng.randrange(6)
        if k == 0:
            sh = s.rectangle(rng.uniform(1, 4), rng.uniform(1, 4))
        elif k == 1:
            sh = s.circle(d=rng.uniform(1, 4))
        elif k == 2:
            sh = s.regular_polygon2d(rng.randrange(3, 9), r=rng.uniform(1, 2.5))
        elif k == 3:
            m = rng.randrange(5, 9)
            pts = [(math.cos(2 * math.pi * i / m) * rng.uniform(1, 3), math.sin(2 * math.pi * i / m) * rng.uniform(1, 3)) for i in range(m)]
            sh = s.polygon2d(pts)
        elif k == 4:
            sh = s.gears.InvoluteGear(rng.randrange(7, 15), rng.uniform(0.4, 0.8))
        else:
            sh = s.rectangle(rng.uniform(2, 5), rng.uniform(0.5, 1.5))
        for _ in range(rng.randrange(0, 3)):
            t = rng.randrange(7)
            if t == 0:
                sh = sh.translated(rng.uniform(-2, 2), rng.uniform(-2, 2))
            elif t == 1:
                sh = sh.rotated(rng.choice([90, 180, 270, rng.uniform(0, 360)]))
            elif t == 2:
                sh = sh.offset(rng.uniform(0.05, 0.3))
            elif t == 3:
                sh = sh.shell(rng.uniform(0.1, 0.3))
            elif t == 4:
                sh = sh.mirrored_x()
            elif t == 5:
                sh = sh.translated(rng.uniform(1, 3), 0).symmetrical_x()
            else:
                sh = sh.scaled(rng.uniform(0.6, 1.5))
        if depth < 2 and rng.random() < 0.5:
            other = shape2d(depth + 1).translated(rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5))
            c = rng.randrange(3)
            sh = sh + other if c == 0 else (sh & other.scaled(2)) if c == 1 else sh - other.scaled(0.5)
        return sh

    def solid():
        sh = shape2d().extruded(rng.uniform(1, 4))
        k = rng.randrange(5)
        if k == 0:
            sh = sh.rotated((0, 0, 1), rng.uniform(0, 360))
        elif k == 1:
            sh = sh.rotated(rng.choice([(1, 0, 0), (0, 1, 0), (0, 0, 1)]), rng.choice([90, 180, 270]))
        elif k == 2:
            sh = sh.rotated((1, 0, 0), 180).rotated((0, 0, 1), rng.uniform(0, 360))
        return sh.translated(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-1, 1))

    out = {}
    for i in range(n):
        sh = solid()
        for _ in range(rng.randrange(0, 3)):
            k = rng.randrange(5)
            if k == 0:
                sh = sh + solid()
            elif k == 1:
                sh = sh - solid().scaled(0.7)
            elif k == 2:
                sh = sh & s.sphere(d=rng.uniform(5, 9)).translated(rng.uniform(-1, 1), rng.uniform(-1, 1), 0)
            elif k == 3:
                sh = sh + s.box(rng.uniform(1, 2)).rotated((1, 2, 3), rng.uniform(0, 90)).translated(rng.uniform(-4, 4), rng.uniform(-4, 4), 0)
            else:
                sh = s.union([sh, solid(), solid()])
        out["colr_%02d" % i] = sh
    return out


def collect():
    scenes = {}

    # --- the five BASELINE.json configs -------------------------------------
    import csg_example
    import menger_sponge
    import airfoil
    import planetary

    scenes["cfg_csg_example"] = csg_example.o
    scenes["cfg_menger_sponge"] = menger_sponge.o
    scenes["cfg_airfoil"] = airfoil.o
    scenes["cfg_planetary"] = planetary.Planetary(11, 60, 13, 41, 18, 53).make_assembly().shape()
    scenes["cfg_synthetic500"] = synthetic_deep_csg(500)
    scenes["cfg_synthetic32"] = synthetic_deep_csg(32)

    # --- tests/data.py shapes (tests/test_dsdf.py) ---------------------------
    import data

    for k, v in sorted(data.shapes_2d.items()):
        scenes["dsdf2d_" + k] = v
    for k, v in sorted(data.shapes_3d.items()):
        scenes["dsdf3d_" + k] = v

    # --- tests/test_mass_properties.py:16-96 ---------------------------------
    hemi = s.sphere(r=2) - s.half_space()
    scenes["mp_unit_box"] = s.box(1)
    scenes["mp_cylinder"] = s.cylinder(h=2, r=4, symmetrical=False)
    scenes["mp_sphere"] = s.sphere(d=2)
    scenes["mp_two_boxes"] = s.box(2).translated(-15, 0, 0) + s.box(2).translated(15, 0, 0)
    scenes["mp_hemisphere"] = hemi
    scenes["mp_translated_sphere"] = s.sphere(d=2).translated(10, 11, 7)
    scenes["mp_translated_and_rotated_hemisphere"] = hemi.translated(2, 0, 0).rotated((1, 0, 0), 90)
    scenes["mp_not_hammer"] = s.box(4).translated(0, 0, 2) + s.box(2, 2, 9).translated(0, 0, -3.5)
    scenes["mp_drunk_box"] = s.box(2, 3, 5).rotated((7, 11, 13), 17)

    # --- tests/test_subdivision.py:110-161 -----------------------------------
    scenes["sub_box10"] = s.box(10)
    res, gs = 0.1, 8
    step = res * (gs - 1)
    scenes["sub_circle"] = s.circle(gs * step - res)

    # --- a few extras exercising ops the above do not ------------------------
    scenes["x_shell_sphere"] = s.sphere(d=6).shell(0.5)
    scenes["x_smooth_isect"] = s.intersection([s.box(4), s.sphere(d=5)], r=0.4)
    scenes["x_scaled_rot"] = s.box(1, 2, 3).scaled(1.7).rotated((1, 1, 0), 33).translated(1, 2, 3)
    scenes["x_twist"] = s.rectangle(2, 1).revolved(4, 90)
    scenes["x_gear3d"] = s.gears.InvoluteGear(13, 1.0).extruded(2)
    scenes["x_repetition"] = codecad.shapes.unsafe.Repetition(s.sphere(d=1), (2, 2, None)) & s.box(7, 7, 2)
    return scenes


def main():
    out = {}
    forests = "--forests" in sys.argv   # the extra file tests/golden/forest_scenes.npz
    columns = "--columns" in sys.argv   # the extra file tests/golden/column_scenes.npz
    scenes = forest_scenes() if forests else column_scenes() if columns else collect()
    for name, shape in scenes.items():
        random.seed(0)
        words = codecad.nodes.make_program(shape)
        assert words.dtype == numpy.float32
        box = shape.bounding_box()
        try:
            fs = float(shape.feature_size())
        except Exception:  # noqa: BLE001 - some shapes have no feature size
            fs = float("nan")
        meta = numpy.array(
            [shape.dimension(), box.a.x, box.a.y, box.a.z, box.b.x, box.b.y, box.b.z, fs],
            dtype=numpy.float64,
        )
        out[name + ".words"] = words
        out[name + ".meta"] = meta
        print("%-46s dim %d  %5d words  bbox %s .. %s" % (name, shape.dimension(), len(words), box.a, box.b))
    path = os.path.join(REPO, "tests", "golden", "forest_scenes.npz" if forests else "column_scenes.npz" if columns else "scenes.npz")
    numpy.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
