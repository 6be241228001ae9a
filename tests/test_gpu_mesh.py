"""Mesh export on the GPU (cc_mesh_blocks behind codecad_b200.rendering.triangular_mesh) against the
CPU restatement of rendering/mesh.py:53-72 (oracle/mc_oracle.py): bit-exact float64 triangles on the
same blocks, plus the reference's own criterion, watertightness (tests/test_mesh.py:12-29; shapes
box(10) and sphere(10), grid sizes 2 / 12 / 16)."""
import numpy as np
import pytest

import oracle
from oracle import host, mc_oracle as mc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import codecad_b200
    from codecad_b200 import _lib
    _lib.init(0)
    return codecad_b200


def _oracle_pieces(scene, grid_size):
    """The reference's loop with the CPU oracle in place of OpenCL and of mcubes."""
    res = scene.feature_size / 2
    if grid_size is None:
        grid_size = 128
    dims, blocks = host.subdivision(scene.words, scene.box_a, scene.box_b, scene.dimension, res, True, grid_size)
    pieces = []
    for box_size, box_corner, box_resolution, *_ in blocks:
        corner32 = np.array(box_corner, np.float64).astype(np.float32)
        field = oracle.grid_eval_pymcubes(scene.words, corner32, np.float32(box_resolution), box_size)
        # the reference hands mcubes the flat buffer viewed with shape max_box_size (mesh.py:20)
        block = field.reshape(-1).reshape(tuple(int(d) for d in dims))
        soup = mc.block_mesh(block, box_corner, box_resolution)
        if len(soup):
            pieces.append(soup)
    return pieces


@pytest.mark.parametrize("grid_size", [2, 12, 16])
@pytest.mark.parametrize("name", ["sub_box10", "dsdf3d_sphere"])
def test_triangles_bit_exact_and_watertight(cb, scenes, name, grid_size):
    from codecad_b200.rendering import triangular_mesh
    s = scenes[name]
    got = [np.asarray(v, np.float64).reshape(-1, 3, 3) for v, _ in triangular_mesh(s.compiled(), grid_size)]
    want = _oracle_pieces(s, grid_size)
    assert len(got) == len(want) and len(got) > 0
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert np.array_equal(g, w)
    soup = np.concatenate(got)
    size = max(b - a for a, b in zip(s.box_a, s.box_b))
    rep = mc.manifold_report(soup, 1e-6 * size)
    assert rep["bad_edges"] == 0, rep                       # the reference's test: watertight
    assert mc.signed_volume(soup) > 0                       # outward normals


def test_mesh_volume_matches_mass_properties(cb, scenes):
    from codecad_b200.rendering import mesh_arrays
    s = scenes["cfg_csg_example"]
    vertices, block, boxes = mesh_arrays(s.compiled(), 32)
    assert len(vertices) > 100 and len(boxes) >= 1 and len(block) == len(vertices)
    # one block at this resolution: shared vertices are bit-identical, no tolerance needed (the scene
    # has samples exactly on the surface, i.e. zero-area triangles that a tolerance would tangle up)
    assert len(boxes) == 1
    rep = mc.manifold_report(vertices, 0)
    assert rep["bad_edges"] == 0, rep
    vol = cb.mass_properties(s.compiled(), 2.0, 32).volume   # the mesh is coarse: feature_size / 2 = 20
    assert mc.signed_volume(vertices) == pytest.approx(vol, rel=0.15)


def test_render_stl(cb, scenes, tmp_path):
    from codecad_b200.rendering import render_stl
    path = tmp_path / "box.stl"
    n = render_stl(scenes["sub_box10"].compiled(), str(path))
    raw = path.read_bytes()
    assert n > 0 and len(raw) == 84 + 50 * n


def test_debug_boxes_and_dimension_check(cb, scenes):
    from codecad_b200.rendering import triangular_mesh
    pieces = list(triangular_mesh(scenes["sub_box10"].compiled(), 4, debug_subdivision_boxes=True))
    assert len(pieces) >= 1 and all(len(v) == 8 and len(t) == 12 for v, t in pieces)
    with pytest.raises(AssertionError):
        list(triangular_mesh(scenes["sub_circle"].compiled()))


@pytest.mark.parametrize("mode", ["pipelined chunks", "one field buffer"])
def test_chunked_paths_give_the_same_triangles(cb, scenes, mode, monkeypatch):
    """cc_mesh_blocks works through the blocks in chunks: with all fields resident the copy of one
    chunk overlaps the emit pass of the next, larger meshes reuse one field buffer.  Tiny budgets
    force both paths on a small mesh; triangles and their order must not change."""
    from codecad_b200 import CompiledScene
    from codecad_b200.rendering import mesh_arrays
    s = scenes["cfg_csg_example"]
    scene = CompiledScene(s.words, 3, s.box_a, s.box_b, 2 * 100.0 / 128, "csg@128")
    want_v, want_b, boxes = mesh_arrays(scene, 16)
    assert len(boxes) > 40 and len(want_v) > 10000
    want_v, want_b = want_v.copy(), want_b.copy()
    block_bytes = 16 ** 3 * 4
    monkeypatch.setenv("CODECAD_B200_MESH_CHUNK_BYTES", str(7 * block_bytes))          # 7 blocks per chunk
    if mode == "one field buffer":
        monkeypatch.setenv("CODECAD_B200_MESH_FIELD_BUDGET", str(20 * block_bytes))    # fewer than the mesh has
    got_v, got_b, _ = mesh_arrays(scene, 16)
    assert np.array_equal(got_b, want_b) and np.array_equal(got_v, want_v)


# ---- pins that do not depend on PyMCubes' exact triangulation (absent from the image) ----------
# Whatever table a marching-cubes implementation uses, its vertices lie on cell edges where the field
# changes sign, so they must be ON the surface to interpolation accuracy, and volume / area of the
# mesh must converge to the solid's as the resolution grows.



@pytest.mark.parametrize("name", ["sub_box10", "dsdf3d_sphere", "cfg_csg_example", "dsdf3d_torus", "mp_drunk_box"])
def test_every_vertex_lies_on_the_surface(cb, scenes, name):
    from codecad_b200 import CompiledScene
    from codecad_b200.rendering import mesh_arrays
    s = scenes[name]
    size = max(b - a for a, b in zip(s.box_a, s.box_b))
    res = size / 96
    scene = CompiledScene(s.words, 3, s.box_a, s.box_b, 2 * res, name + "@96")
    vertices, _, boxes = mesh_arrays(scene, 32)
    pts = np.asarray(vertices, np.float64).reshape(-1, 3).copy()
    assert len(pts) > 1000
    # The reference's post-transform (rendering/mesh.py:68-72) undoes the y flip of grid_eval_pymcubes
    # by negating the flipped row index, which leaves every block — hence the whole mesh — translated
    # by -(ny - 1) * resolution along y.  Reproduced faithfully (STL consumers see the reference's
    # coordinates); undone here to compare with the field.
    pts[:, 1] += (int(boxes.dims[1]) - 1) * res
    d = cb.evaluate_points(scene, pts.astype(np.float32))[:, 3]
    # linear interpolation along a cell edge of a 1-Lipschitz field: the error is second order in the
    # cell size where the field is smooth and at most one cell near creases
    assert np.abs(d).max() <= 1.0 * res, np.abs(d).max() / res
    assert np.mean(np.abs(d)) <= 0.05 * res, np.mean(np.abs(d)) / res


def _area(soup):
    a, b, c = soup[:, 0], soup[:, 1], soup[:, 2]
    return 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()


def test_volume_and_area_converge_to_the_analytic_solid(cb, scenes):
    from codecad_b200 import CompiledScene
    from codecad_b200.rendering import mesh_arrays
    # box(10): volume 1000, area 600.  sphere: radius from its bounding box.
    for name, volume, area in (("sub_box10", 1000.0, 600.0), ("dsdf3d_sphere", None, None)):
        s = scenes[name]
        size = max(b - a for a, b in zip(s.box_a, s.box_b))
        if volume is None:
            r = size / 2
            volume, area = 4 / 3 * np.pi * r ** 3, 4 * np.pi * r ** 2
        errs = []
        for n in (24, 48, 96, 192):
            scene = CompiledScene(s.words, 3, s.box_a, s.box_b, 2 * size / n, "%s@%d" % (name, n))
            soup = np.asarray(mesh_arrays(scene, 32)[0], np.float64).reshape(-1, 3, 3)   # (translated along y: no effect here)
            errs.append((abs(mc.signed_volume(soup) - volume) / volume, abs(_area(soup) - area) / area))
        v_err, a_err = zip(*errs)
        assert v_err[-1] <= 2e-3 and a_err[-1] <= 2e-2, (name, errs)
        assert v_err[-1] <= v_err[0] and a_err[-1] <= a_err[0] * 1.01, (name, errs)
