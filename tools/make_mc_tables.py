#!/usr/bin/env python3
"""Generate the marching-cubes case table used by csrc/cc_mesh.cu and oracle/mc_oracle.py.

Conventions are those of PyMCubes 0.0.6 / P. Bourke's polygonise (what the reference calls through
mcubes.marching_cubes, rendering/mesh.py:63; the package is not vendored, so the table cannot be
copied or checked against it):

  corners  v0=(0,0,0) v1=(1,0,0) v2=(1,1,0) v3=(0,1,0) v4=(0,0,1) v5=(1,0,1) v6=(1,1,1) v7=(0,1,1)
  edges    e0=v0v1 e1=v1v2 e2=v2v3 e3=v3v0 e4=v4v5 e5=v5v6 e6=v6v7 e7=v7v4 e8=v0v4 e9=v1v5 e10=v2v6 e11=v3v7
  case     bit m set  <=>  value(v_m) <= isovalue  ("inside")
  winding  triangle normals point towards the inside corners (case 1 -> triangle e0,e8,e3)

The table itself is GENERATED, not transcribed: for every case the crossing edges are linked face by
face into closed loops and each loop is triangulated with diagonals that run through the cell's
interior (never inside a cube face, see triangulate()).  On an ambiguous face (two diagonal inside
corners) every inside corner is cut off separately.  That rule depends only on the four corner states
of the face, so neighbouring cells always agree on the shared face and the surface is watertight by
construction (the classic Bourke/Bloyd table agrees with it on the 15 base cases but resolves some
complementary ambiguous cases the other way, which is where its well-known holes come from).

    python tools/make_mc_tables.py > codecad_b200/csrc/cc_mc_table.h
"""
import sys

CORNERS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
FACES = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (3, 2, 6, 7), (0, 3, 7, 4), (1, 2, 6, 5)]
# Interpolation direction.  PyMCubes creates every vertex once, in the first cell that touches its
# edge, and shares it by index; this port emits a triangle soup, so every cell recomputes the vertex.
# Interpolating always from the corner with the lower coordinate makes the cells sharing an edge
# compute bit-identical vertices (e2, e3, e6, e7 run high -> low in Bourke's numbering).
EDGES_LOW_HIGH = [(a, b) if CORNERS[a] < CORNERS[b] else (b, a) for a, b in EDGES]
EDGE_OF = {}
for _i, (_a, _b) in enumerate(EDGES):
    EDGE_OF[(_a, _b)] = EDGE_OF[(_b, _a)] = _i


def case_segments(case):
    """Undirected segments (cube edge, cube edge) drawn on the six faces."""
    inside = [(case >> m) & 1 for m in range(8)]
    segs = []
    for face in FACES:
        crossing = []
        for k in range(4):
            a, b = face[k], face[(k + 1) % 4]
            if inside[a] != inside[b]:
                crossing.append(EDGE_OF[(a, b)])
        if len(crossing) == 2:
            segs.append((crossing[0], crossing[1]))
        elif len(crossing) == 4:
            for k in range(4):  # ambiguous face: cut off every inside corner separately
                if inside[face[k]]:
                    segs.append((EDGE_OF[(face[(k - 1) % 4], face[k])], EDGE_OF[(face[k], face[(k + 1) % 4])]))
    return inside, segs


def case_triangles(case):
    inside, segs = case_segments(case)
    link = {}
    for a, b in segs:
        link.setdefault(a, []).append(b)
        link.setdefault(b, []).append(a)
    assert all(len(v) == 2 for v in link.values()), (case, link)
    mid = [tuple((CORNERS[a][i] + CORNERS[b][i]) / 2 for i in range(3)) for a, b in EDGES]
    tris, seen = [], set()
    for start in sorted(link):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = link[cur]
            nxt = b if a == prev else a
            if prev is None:
                nxt = a
            if nxt == start:
                break
            assert nxt not in seen, (case, loop, nxt)
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        assert len(loop) >= 3, (case, loop)
        # orientation: the Newell normal must point towards the inside corners
        nrm = [0.0, 0.0, 0.0]
        for i in range(len(loop)):
            p, q = mid[loop[i]], mid[loop[(i + 1) % len(loop)]]
            nrm[0] += (p[1] - q[1]) * (p[2] + q[2])
            nrm[1] += (p[2] - q[2]) * (p[0] + q[0])
            nrm[2] += (p[0] - q[0]) * (p[1] + q[1])
        score = 0.0
        for e in loop:
            a, b = EDGES[e]
            if not inside[a]:
                a, b = b, a
            score += sum(n * (ca - cb) for n, ca, cb in zip(nrm, CORNERS[a], CORNERS[b]))
        assert abs(score) > 1e-9, (case, loop)
        if score < 0:
            loop = [loop[0]] + loop[:0:-1]
        tris.extend(triangulate(loop))
    return tris


def on_common_face(e1, e2):
    """Do the two cube edges lie on a common cube face?"""
    for face in FACES:
        fe = {EDGE_OF[(face[k], face[(k + 1) % 4])] for k in range(4)}
        if e1 in fe and e2 in fe:
            return True
    return False


def triangulations(poly):
    """All triangulations of a convex-position polygon (list of vertex ids), as lists of triangles
    that keep the polygon's orientation."""
    if len(poly) < 3:
        return [[]]
    if len(poly) == 3:
        return [[tuple(poly)]]
    out = []
    a, b = poly[0], poly[-1]
    for k in range(1, len(poly) - 1):  # triangle (a, poly[k], b) splits the polygon
        left = triangulations(poly[:k + 1])
        right = triangulations(poly[k:])
        for lt in left:
            for rt in right:
                out.append(lt + [(a, poly[k], b)] + rt)
    return out


def triangulate(loop):
    """Triangulate a loop of cube edges.  A diagonal between two vertices that lie on a common cube
    face would lie IN that face, where the neighbouring cell may draw the same diagonal: four
    triangles would then share one edge.  Pick the first triangulation (deterministic enumeration
    order) whose diagonals all run through the cell's interior."""
    best, best_bad = None, None
    sides = {frozenset((loop[i], loop[(i + 1) % len(loop)])) for i in range(len(loop))}
    for tri_set in triangulations(list(loop)):
        bad = 0
        for t in tri_set:
            for e in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if frozenset(e) not in sides and on_common_face(*e):
                    bad += 1
        if best is None or bad < best_bad:
            best, best_bad = tri_set, bad
        if bad == 0:
            break
    assert best_bad == 0, (loop, best_bad)
    return list(best)


def build_table():
    table = [case_triangles(c) for c in range(256)]
    assert max(len(t) for t in table) <= 5
    assert table[0] == [] and table[255] == []
    assert table[1] in ([(0, 8, 3)], [(8, 3, 0)], [(3, 0, 8)]), table[1]   # Bourke's case 1, same winding
    return table


def main():
    table = build_table()
    out = sys.stdout
    out.write("// GENERATED by tools/make_mc_tables.py - do not edit.  Conventions: see that script.\n")
    out.write("#ifndef CC_MC_TABLE_H\n#define CC_MC_TABLE_H\n")
    out.write("// number of triangles per case\nstatic const unsigned char cc_mc_count[256] = {\n")
    for r in range(0, 256, 32):
        out.write("    " + ", ".join(str(len(table[c])) for c in range(r, r + 32)) + ",\n")
    out.write("};\n// cube edges (3 per triangle, up to 5 triangles), 255 = end\nstatic const unsigned char cc_mc_tri[256][16] = {\n")
    for c in range(256):
        flat = [e for t in table[c] for e in t]
        flat += [255] * (16 - len(flat))
        out.write("    {" + ", ".join("%d" % e for e in flat) + "},\n")
    out.write("};\n// the two corners of every cube edge, lower coordinate first (interpolation direction)\n"
              "static const unsigned char cc_mc_edge_corner[12][2] = {\n    ")
    out.write(", ".join("{%d, %d}" % e for e in EDGES_LOW_HIGH) + "\n};\n")
    out.write("// corner offsets (i, j, k)\nstatic const unsigned char cc_mc_corner[8][3] = {\n    ")
    out.write(", ".join("{%d, %d, %d}" % c for c in CORNERS) + "\n};\n#endif\n")


if __name__ == "__main__":
    main()
