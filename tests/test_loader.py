"""Program loader (wire format -> device microcode), host-only through cc_program_decode:
no GPU needed.  Covers validation/error behaviour, liveness renaming, store folding,
primitive fusion and the static flop count."""
import collections

import numpy as np
import pytest

from codecad_b200 import _lib
from codecad_b200.opcodes import OPCODE, REGISTER_COUNT, disassemble, disassemble_microcode, SLOT_NONE
from scenes import ALL_NAMES


def ins(name, reg=0, *params):
    return [float(OPCODE[name] * REGISTER_COUNT + reg)] + [float(p) for p in params]


IDENT = (0, 0, 0, 1, 0, 0, 0)


@pytest.mark.parametrize("name", ALL_NAMES)
def test_every_fixture_decodes(scenes, name):
    s = scenes[name]
    info, code = _lib.decode_program(s.words)
    wire = list(disassemble(s.words))
    assert info.n_instructions == len(wire)
    assert info.n_words == len(s.words)
    assert len(code) == info.n_micro_words and len(code) % 4 == 0
    ops = list(disassemble_microcode(code))
    assert len(ops) == info.n_micro_ops and ops[-1][1] == "RETURN"
    # polygon edge tables (6 words per vertex) live after RETURN
    assert ops[-1][0] + 4 <= len(code)
    # every _store is folded away or dropped; slots are dense and never exceed the wire count
    assert info.n_slots <= max(1, info.n_wire_registers)
    used = {o[2] for o in ops if o[2] != SLOT_NONE} | {o[3] for o in ops if o[3] != SLOT_NONE}
    assert used <= set(range(info.n_slots))
    # a slot is written before it is read
    written = set()
    for _, opname, src, dst, _ in ops:
        if src != SLOT_NONE:
            assert src in written, "slot %d read before written" % src
        if dst != SLOT_NONE:
            written.add(dst)
    assert info.flops_min <= info.flops_max


def test_config_scene_statistics(scenes):
    """SURVEY.md 8 scene table: instruction counts and the (much smaller) live sets."""
    expect = {"cfg_csg_example": 28, "cfg_menger_sponge": 115, "cfg_airfoil": 119, "cfg_planetary": 467,
              "cfg_synthetic500": 3999}
    for name, n in expect.items():
        info, _ = _lib.decode_program(scenes[name].words)
        assert info.n_instructions == n
        assert info.n_slots <= 10
    info, code = _lib.decode_program(scenes["cfg_synthetic500"].words)
    c = collections.Counter(o[1] for o in disassemble_microcode(code))
    assert c["PRIM_RECT"] == 500 and c["UNION_R"] == 499 and info.n_micro_ops == 1000
    assert info.n_wire_registers >= 500 and info.n_slots < 12   # the reference scheduler leaks registers


def test_flop_count_matches_survey(scenes):
    """static algorithmic flop/point, SURVEY.md 8(d): csg 264-339, menger 1714-1984,
    planetary 6985-7600, synthetic 52 491-73 479"""
    want = {"cfg_csg_example": (264, 339), "cfg_menger_sponge": (1714, 1984), "cfg_planetary": (6985, 7600),
            "cfg_synthetic500": (52491, 73479)}
    for name, (lo, hi) in want.items():
        info, _ = _lib.decode_program(scenes[name].words)
        assert (info.flops_min, info.flops_max) == (lo, hi)


def test_box_is_one_fused_primitive():
    # box(1): SURVEY.md appendix A worked example
    words = ins("initial_transformation_to", 0, *IDENT) + ins("_store", 0) + ins("rectangle", 0, .5, .5) \
        + ins("extrusion", 0, .5) + ins("_return")
    info, code = _lib.decode_program(words)
    ops = [o[1] for o in disassemble_microcode(code)]
    assert ops == ["PRIM_RECT_M", "RETURN"] and info.n_fused == 1 and info.n_slots == 0   # identity matrices: masked variant
    f = code.view(np.float32)
    assert list(f[1:10]) == [1, 0, 0, 0, 1, 0, 0, 0, 1]            # rotation matrix of the unit quaternion
    assert list(f[13:17]) == [.5, .5, .5, 0]                        # hw, hh, h, offset 0
    assert list(f[17:27]) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 1]         # identity transformation_from, scale 1


def test_point_with_two_readers_is_not_fused():
    words = ins("initial_transformation_to", 0, *IDENT) + ins("_store", 3) + ins("circle", 0, 2) \
        + ins("extrusion", 3, 1) + ins("_store", 5) + ins("_load", 3) + ins("sphere", 0, 1) \
        + ins("union", 5, -1) + ins("_return")
    info, code = _lib.decode_program(words)
    ops = list(disassemble_microcode(code))
    assert [o[1] for o in ops] == ["T_INIT_M", "CIRCLE", "EXTRUSION", "LOAD", "SPHERE", "UNION", "RETURN"]
    assert info.n_fused == 0 and info.n_slots == 2
    assert ops[0][3] == 0 and ops[2][2] == 0 and ops[2][3] == 1 and ops[3][2] == 0 and ops[5][2] == 1


def test_dead_store_is_dropped_and_registers_are_recycled():
    words = ins("initial_transformation_to", 0, *IDENT)
    for r in range(40):                       # 40 wire registers, one live at a time
        words += ins("_store", r) + ins("sphere", 0, 1) + ins("union", r, -1)
    words += ins("_store", 77) + ins("_return")   # never read
    info, code = _lib.decode_program(words)
    assert info.n_wire_registers == 78 and info.n_slots == 1
    assert all(o[3] in (0, SLOT_NONE) for o in disassemble_microcode(code))


def test_rounded_versus_sharp_combinators():
    base = ins("initial_transformation_to", 0, *IDENT) + ins("_store", 0) + ins("sphere", 0, 1)
    for name, sharp, rnd in (("union", "UNION", "UNION_R"), ("intersection", "ISECT", "ISECT_R"),
                             ("subtraction", "SUB", "SUB_R")):
        for r, want in ((-1.0, sharp), (0.0, rnd), (0.5, rnd)):
            _, code = _lib.decode_program(base + ins(name, 0, r) + ins("_return"))
            assert [o[1] for o in disassemble_microcode(code)][-2] == want


@pytest.mark.parametrize("words,fragment", [
    ([], "without _return"),
    (ins("initial_transformation_to", 0, *IDENT), "without _return"),
    (ins("initial_transformation_to", 0, 0, 0), "truncated"),
    ([float(29 * 512)], "invalid instruction word"),
    ([-1.0], "invalid instruction word"),
    ([5632.5], "invalid instruction word"),
    ([float("nan")], "invalid instruction word"),
    (ins("sphere", 0, 1) + ins("_return"), "must start with"),
    (ins("initial_transformation_to", 0, *IDENT) + ins("_load", 4) + ins("_return"), "before any _store"),
    (ins("initial_transformation_to", 0, *IDENT) + ins("union", 2, -1) + ins("_return"), "before any _store"),
    (ins("initial_transformation_to", 0, *IDENT) + ins("polygon2d", 0, 0) + ins("_return"), "polygon2d"),
    (ins("initial_transformation_to", 0, *IDENT) + ins("polygon2d", 0, 3, 0, 0, 1, 0), "truncated"),
])
def test_malformed_programs_are_rejected(words, fragment):
    with pytest.raises(_lib.CodecadB200Error) as e:
        _lib.decode_program(np.array(words, np.float32))
    assert fragment in str(e.value)


def test_words_after_return_are_ignored():
    words = ins("initial_transformation_to", 0, *IDENT) + ins("sphere", 0, 1) + ins("_return") + [123.0, 456.0]
    info, _ = _lib.decode_program(words)
    assert info.n_words == len(words) - 2 and info.n_instructions == 3


def test_matrix_zero_masks_and_identity_elision():
    """cc-arith omits zero matrix coefficients: the loader tags such transforms "_M" with a mask
    word, keeps full matrices on the plain micro-ops and drops identity transformation_from ops."""
    import math
    q_rot = (0.0, 0.0, math.sin(0.3), math.cos(0.3))            # rotation about z: zeros in the matrix
    q_gen = (0.1825742, 0.3651484, 0.5477226, 0.7302967)        # general rotation: no zero
    words = ins("initial_transformation_to", 0, *q_rot, 1.0, 2.0, 3.0) + ins("sphere", 0, 1) \
        + ins("transformation_from", 0, 0.0, 0.0, 0.0, 1.0) + ins("_store", 1) \
        + ins("initial_transformation_to", 0, *q_gen, 0.0, 0.0, 0.0) + ins("sphere", 0, 2) \
        + ins("transformation_from", 0, *q_gen) + ins("union", 1, -1) + ins("_return")
    info, code = _lib.decode_program(words)
    ops = list(disassemble_microcode(code))
    assert [o[1] for o in ops] == ["T_INIT_M", "SPHERE", "T_INIT", "SPHERE", "T_FROM", "UNION", "RETURN"]
    assert ops[1][3] != 0x1FF                                    # the _store folded into SPHERE (identity T_FROM elided)
    mask = int(code[ops[0][0] + 13])
    m = code.view(np.float32)[ops[0][0] + 1: ops[0][0] + 10]
    assert mask == sum(1 << i for i in range(9) if m[i] != 0) == 0x11B


def test_union_forests_are_recognised_and_nothing_else(scenes):
    """cc_program.cpp analyse_forest: one tree of unions over fused primitives -> forest tables."""
    from scenes import FOREST_NAMES
    from codecad_b200 import _lib
    want = {"cfg_synthetic500": (500, 9), "cfg_synthetic32": (32, 5), "forest_dense64": (64, None)}
    for name, s in scenes.items():
        info, _ = _lib.decode_program(s.words)
        if name in FOREST_NAMES:
            assert info.n_forest_leaves == info.n_fused > 0, name
            assert 1 <= info.forest_depth <= info.n_forest_leaves
            if name in want:
                assert info.n_forest_leaves == want[name][0]
                assert want[name][1] in (None, info.forest_depth)
        else:
            assert info.n_forest_leaves == 0 and info.forest_depth == 0, name


def test_assemblies_are_cut_into_parts(scenes):
    """cc_program.cpp analyse_parts: sharp unions over self-contained sub-programs, each with a Lipschitz
    bound unless it contains an op that has none."""
    from codecad_b200 import _lib
    want = {"cfg_planetary": (7, 7), "cfg_menger_sponge": (2, 1), "dsdf3d_mirror_3d": (5, 5), "dsdf2d_mirror_2d": (4, 4),
            "dsdf2d_rotated_pattern_2d": (3, 3)}
    for name, s in scenes.items():
        info, _ = _lib.decode_program(s.words)
        if name in want:
            assert (info.n_parts, info.n_parts_bounded) == want[name], name
        assert info.n_parts_bounded <= info.n_parts <= 32
        if name in ("cfg_csg_example", "cfg_airfoil", "mp_sphere", "x_smooth_isect"):
            assert info.n_parts == 0, name        # one component, or a rounded combinator at the root


def test_profiles_under_extrusions_are_column_invariant(scenes):
    """cc_program.cpp analyse_columns: which micro-ops cannot see the grid coordinate along the best axis, and
    that the code generator turns the split into the column kernels (source only: no device needed)."""
    from codecad_b200 import _lib
    want_axis = {"cfg_planetary": 2, "x_gear3d": 2, "dsdf2d_gear": 2}
    for name, s in scenes.items():
        info, _ = _lib.decode_program(s.words)
        assert info.column_invariant_percent == 0 or 25 <= info.column_invariant_percent <= 100, name
        assert info.column_axis in (0, 1, 2)
        if name in want_axis:
            assert info.column_invariant_percent >= 40 and info.column_axis == want_axis[name], name
        if s.dimension == 2 and info.n_micro_ops >= 3:
            # a 2-D scene never reads z (scenes whose transformations tilt the plane aside)
            assert info.column_invariant_percent in (0, 100) or info.column_axis != 2, name
        if name in ("cfg_csg_example", "mp_sphere"):
            assert info.column_invariant_percent == 0, name   # spheres and boxes in general position: nothing to hoist
    assert _lib.decode_program(scenes["cfg_airfoil"].words)[0].column_axis != 2     # the wing's span is not the grid's z
    src = _lib.specialize_source(scenes["cfg_planetary"].words, 2, compile=False, sink_mask=256)
    src = src[0] if isinstance(src, tuple) else src
    for kernel in ("cc_jit_columns_centers", "cc_jit_columns_profiles", "cc_jit_columns(", "cc_jit_columns_full"):
        assert kernel in src, kernel
    ahead = src[src.index("struct SceneAhead"):src.index("struct SceneEval")]
    loop = src[src.index("struct SceneEval"):src.index("struct SceneTile")]
    assert ahead.count("cc_op_gear") == 9 and loop.count("cc_op_gear") == 0        # the nine involute gears leave the per-cell body
    assert loop.count("cc_extrusion_n") >= 12 and "cc_col_load" in loop and "cc_col_store" in ahead
    assert "cc_same_bits" in ahead                                                 # the half turns' residue rows are checked per column
    src = _lib.specialize_source(scenes["x_gear3d"].words, 2, compile=False, sink_mask=256)
    src = src[0] if isinstance(src, tuple) else src
    assert "cc_jit_columns_centers" not in src                                     # no parts, no brick centres


def test_tile_units_for_the_hierarchy_sinks(scenes):
    """The hierarchy sinks (blocks x linear tiles) of a program with parts or a column split have a unit each: tile-centre
    pass and / or column pass + the tile kernel (source only)."""
    from codecad_b200 import _lib
    TILES_PYMCUBES, TILES_CLASSIFY, TILES_MASS = 1 << 9, 1 << 10, 1 << 11
    src = _lib.specialize_source(scenes["dsdf3d_mirror_3d"].words, 2, compile=False, sink_mask=TILES_CLASSIFY)
    src = src[0] if isinstance(src, tuple) else src
    assert "cc_jit_tile_centers" in src and "cc_jit_tile_classify" in src and "cc_jit_tile_profiles" not in src   # parts, no columns
    src = _lib.specialize_source(scenes["cfg_airfoil"].words, 2, compile=False, sink_mask=TILES_MASS)
    src = src[0] if isinstance(src, tuple) else src
    assert "cc_jit_tile_profiles" in src and "cc_jit_tile_mass" in src and "cc_jit_tile_centers" not in src        # columns, no parts
    src = _lib.specialize_source(scenes["cfg_planetary"].words, 2, compile=False, sink_mask=TILES_PYMCUBES)
    src = src[0] if isinstance(src, tuple) else src
    for kernel in ("cc_jit_tile_centers", "cc_jit_tile_profiles", "cc_jit_tile_pymcubes"):
        assert kernel in src, kernel
    assert "a.part_masks[tile]" in src
    # the dense-grid units carry no tile kernels (compile time)
    src = _lib.specialize_source(scenes["cfg_planetary"].words, 2, compile=False, sink_mask=128)
    src = src[0] if isinstance(src, tuple) else src
    assert "cc_jit_parts(" in src and "cc_jit_part_centers" in src and "cc_jit_tile" not in src


def test_column_split_is_self_consistent():
    """cc_program.cpp analyse_columns on every fixture: what the per-cell body reads is either computed per cell or handed
    over by the column pass; nothing the column pass runs can see the axis, except ops that feed it in their invariant
    components (initial transforms) and the cut primitives; the column pass is closed under its inputs."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path[:0] = [%r, %r]\n"
            "from codecad_b200 import _lib\n"
            "from scenes import load_scenes\n"
            "for n, s in sorted(load_scenes().items()):\n"
            "    sys.stderr.write('scene %%s\\n' %% n); sys.stderr.flush()\n"
            "    _lib.decode_program(s.words)\n") % (root, os.path.join(root, "tests"))
    env = dict(os.environ, CODECAD_B200_COLUMNS_DEBUG="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    pat = re.compile(r"op\s+(\d+)\s+mop\s+(\d+)\s+in_l\s+(-?\d+) in_s\s+(-?\d+)\s+dep (\w+)\s+phase (\d+)\s+save (\d+) restore (-?\d+) cost (\d+)")
    scenes_seen, tables = 0, 0
    table = []

    def check(rows):
        ops = {r[0]: r for r in rows}
        # ops of the column pass that reach, through column-pass ops, one whose value cannot see the axis
        feeds_invariant = {r[0] for r in rows if r[5] & 1 and r[4] == 0}
        for r in reversed(rows):                             # (inputs precede their readers)
            if r[0] in feeds_invariant:
                feeds_invariant.update(p for p in (r[2], r[3]) if p >= 0)
        for i, mop, in_l, in_s, dep, phase, save, restore, cost in rows:
            if phase & 2:                                   # runs per cell
                assert dep != 0, "an invariant op in the per-cell body"
                for p in (in_l, in_s):
                    if p < 0:
                        continue
                    if ops[p][5] & 2:
                        continue                            # computed per cell as well
                    assert ops[p][4] == 0, "the per-cell body reads a dependent value it does not compute"
                    assert ops[p][5] & 1, "... or an invariant one the column pass does not compute"
                if in_l >= 0 and not ops[in_l][5] & 2:
                    assert restore == in_l and ops[in_l][6] == 1, "running value not handed over"
                else:
                    assert restore == -1
            if phase & 1:                                   # runs in the column pass
                for p in (in_l, in_s):
                    assert p < 0 or ops[p][5] & 1, "the column pass is not closed under its inputs"
                if dep != 0 and mop not in (32, 33, 37, 38):     # (a cut primitive is there for its profile half)
                    assert i in feeds_invariant, "a dependent op in the column pass that feeds nothing invariant"
            if save:
                assert phase & 1 and not phase & 2 and dep == 0

    for line in out.stderr.splitlines():
        if line.startswith("scene "):
            scenes_seen += 1
            continue
        m = pat.match(line)
        if not m:
            continue
        row = (int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5), 16), int(m.group(6)), int(m.group(7)),
               int(m.group(8)), int(m.group(9)))
        if row[0] == 0 and table:
            check(table)
            tables += 1
            table = []
        table.append(row)
    if table:
        check(table)
        tables += 1
    assert scenes_seen > 100 and tables >= scenes_seen          # (one table per axis and scene that got as far as the split)


@pytest.mark.parametrize("name,sink", [("col_star_x", 256), ("col_star_half_turn", 256), ("col_hole_grid", 256), ("col_hole_grid", 2048),
                                       ("dsdf3d_mirror_3d", 128), ("dsdf3d_mirror_3d", 1024), ("colr_08", 512), ("cfg_airfoil", 2048)])
def test_brick_and_tile_units_compile(scenes, name, sink):
    """The generated source of the part-culling (128), column (256) and tile units (512 / 1024 / 2048: PyMCubes / classify /
    mass) goes through NVRTC for sm_100a here, without a device: columns along x and y, residue rows with the per-column
    check, parts with and without a column split, a cut primitive."""
    from codecad_b200 import _lib
    try:
        src, cubin_bytes = _lib.specialize_source(scenes[name].words, 2, compile=True, sink_mask=sink)
    except _lib.CodecadB200Error as exc:
        if "NVRTC" in str(exc) and "not found" in str(exc):
            pytest.skip("no libnvrtc in this environment")
        raise
    assert cubin_bytes > 10000 and '__global__' in src
