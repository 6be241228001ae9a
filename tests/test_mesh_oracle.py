"""Mesh export without a GPU: the generated marching-cubes table, the CPU restatement of the
reference's mesh step (oracle/mc_oracle.py, rendering/mesh.py:53-72) and the STL writer.
The reference's own test of this step is watertightness (tests/test_mesh.py:12-29)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import make_mc_tables as gen  # noqa: E402
from oracle import mc_oracle as mc  # noqa: E402


def test_table_matches_generated_header():
    """csrc/cc_mc_table.h is exactly what the generator prints."""
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gen.main()
    assert buf.getvalue() == open(os.path.join(ROOT, "codecad_b200", "csrc", "cc_mc_table.h")).read()


def test_table_case_properties():
    table = gen.build_table()
    assert table[0] == [] and table[255] == []
    assert max(len(t) for t in table) == 5
    for case in range(256):
        inside = [(case >> m) & 1 for m in range(8)]
        used = {e for t in table[case] for e in t}
        crossing = {i for i, (a, b) in enumerate(gen.EDGES) if inside[a] != inside[b]}
        assert used == crossing, case                    # every crossed edge is used, no other
        # each triangle edge that is not shared inside the cell lies on a cube face
        directed = {}
        for a, b, c in table[case]:
            for e in ((a, b), (b, c), (c, a)):
                directed[e] = directed.get(e, 0) + 1
        assert all(n == 1 for n in directed.values()), case
        open_edges = [e for e in directed if (e[1], e[0]) not in directed]
        _, segs = gen.case_segments(case)
        face_segments = {frozenset(s) for s in segs}
        assert {frozenset(e) for e in open_edges} == face_segments, case


def test_face_rule_is_local():
    """The segments drawn on a face depend only on that face's four corner states, so two cells
    sharing a face always draw the same segments there (watertight by construction)."""
    for f, face in enumerate(gen.FACES):
        seen = {}
        for case in range(256):
            inside, segs = gen.case_segments(case)
            face_edges = {gen.EDGE_OF[(face[k], face[(k + 1) % 4])] for k in range(4)}
            on_face = frozenset(frozenset(s) for s in segs if set(s) <= face_edges and _same_face(s, face))
            key = tuple(inside[c] for c in face)
            assert seen.setdefault(key, on_face) == on_face, (f, case)


def _same_face(seg, face):
    # both cube edges of the segment belong to this face (two edges can share two faces only if equal)
    edges = [{face[k], face[(k + 1) % 4]} for k in range(4)]
    return all(set(gen.EDGES[e]) in edges for e in seg)


@pytest.mark.parametrize("n", [7, 12])
def test_sphere_is_closed_and_oriented(n):
    g = np.linspace(-1.3, 1.3, n)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    field = (np.sqrt(x * x + y * y + z * z) - 1.0).astype(np.float32)
    soup = mc.marching_cubes(field)
    rep = mc.manifold_report(soup, 1e-9)
    assert rep["bad_edges"] == 0 and rep["triangles"] > 0
    h = 2.6 / (n - 1)
    vol = -mc.signed_volume(soup) * h ** 3      # table winding: normals point inside (Bourke)
    assert vol == pytest.approx(4 / 3 * np.pi, rel=0.12)
    # after the reference's post-transform the normals point outwards
    world = mc.block_mesh(field, (-1.3, 1.3, -1.3), h)
    assert mc.signed_volume(world) == pytest.approx(vol, rel=1e-9)
    # vertices lie on cell edges, between the two samples
    v = soup.reshape(-1, 3)
    on_grid = np.isclose(v, np.round(v), atol=0).sum(axis=1)
    assert (on_grid >= 2).all()


def test_random_field_is_closed():
    rng = np.random.default_rng(7)
    for _ in range(5):
        f = rng.normal(size=(9, 8, 10)).astype(np.float32)
        f[0] = f[-1] = 1
        f[:, 0] = f[:, -1] = 1
        f[:, :, 0] = f[:, :, -1] = 1
        soup = mc.marching_cubes(f)
        assert mc.manifold_report(soup, 1e-9)["bad_edges"] == 0


def test_empty_and_full_blocks():
    assert len(mc.marching_cubes(np.ones((4, 4, 4), np.float32))) == 0
    assert len(mc.marching_cubes(-np.ones((4, 4, 4), np.float32))) == 0
    assert len(mc.block_mesh(np.ones((3, 3, 3), np.float32), (0, 0, 0), 1.0)) == 0


def test_binary_stl_round_trip(tmp_path):
    from codecad_b200.rendering.stl_renderer import write_binary_stl
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1], [0, 1, 1], [1, 0, 1]]], dtype=np.float64)
    path = tmp_path / "t.stl"
    assert write_binary_stl(str(path), tri) == 2
    raw = path.read_bytes()
    assert len(raw) == 80 + 4 + 2 * 50
    assert np.frombuffer(raw[80:84], "<u4")[0] == 2
    rec = np.frombuffer(raw[84:], dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    assert np.array_equal(rec["v"], tri.astype(np.float32))
    assert np.array_equal(rec["n"], [[0, 0, 1], [0, 0, -1]])
