"""CPU restatement of the mesh-export step — TEST INFRASTRUCTURE ONLY (never imported by
codecad_b200/).  Follows /root/reference/codecad/rendering/mesh.py:53-72 around
mcubes.marching_cubes(block, 0).

PARITY UNPINNED for the triangulation itself: mcubes is PyMCubes 0.0.6 (requirements.txt:11), a
third-party package that is neither in /root/reference nor installed here, and the reference holds
no golden mesh.  What is restated is its published algorithm (P. Bourke's polygonise conventions:
corner / edge numbering, inside = value <= isovalue, cells visited in (i, j, k) order, vertex =
c1 + (iso - f1) * (c2 - c1) / (f2 - f1) in double precision) with the case table of
tools/make_mc_tables.py, which differs from PyMCubes' transcribed table on complementary ambiguous
cases (see that script).  The reference's own test for this step is watertightness
(tests/test_mesh.py:12-29), which tests/test_mesh*.py check with a local manifold test.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import make_mc_tables as _t  # noqa: E402

TABLE = _t.build_table()
EDGES = _t.EDGES_LOW_HIGH   # interpolation always from the lower-coordinate corner
CORNERS = _t.CORNERS


def marching_cubes(block):
    """block: float32 array [d0][d1][d2].  Returns a soup float64 [t][3][3] in index coordinates,
    cells in (i, j, k) order, triangles in table order, vertices in table order."""
    block = np.asarray(block, dtype=np.float32)
    d0, d1, d2 = block.shape
    out = []
    inside = block <= np.float32(0.0)
    for i in range(d0 - 1):
        for j in range(d1 - 1):
            for k in range(d2 - 1):
                case = 0
                v = []
                for m, (a, b, c) in enumerate(CORNERS):
                    v.append(float(block[i + a, j + b, k + c]))
                    if inside[i + a, j + b, k + c]:
                        case |= 1 << m
                for tri in TABLE[case]:
                    pts = []
                    for e in tri:
                        ca, cb = EDGES[e]
                        p = []
                        for ax, base in enumerate((i, j, k)):
                            a1 = float(base + CORNERS[ca][ax])
                            a2 = float(base + CORNERS[cb][ax])
                            if a1 != a2:
                                f1, f2 = v[ca], v[cb]
                                p.append(a1 + (0.0 - f1) * (a2 - a1) / (f2 - f1))
                            else:
                                p.append(a1)
                        pts.append(p)
                    out.append(pts)
    return np.array(out, dtype=np.float64).reshape(-1, 3, 3)


def block_mesh(block, box_corner, box_resolution):
    """rendering/mesh.py:63-72 for one block: soup in world coordinates, winding flipped."""
    soup = marching_cubes(block)
    if not len(soup):
        return soup
    v = soup.reshape(-1, 3).copy()
    v[:, [0, 1]] = v[:, [1, 0]]
    v[:, 1] *= -1
    v *= float(box_resolution)
    v += np.array(box_corner, dtype=np.float64)
    soup = v.reshape(-1, 3, 3)
    return soup[:, [1, 0, 2], :]


def manifold_report(soup, tol):
    """Merge vertices closer than `tol` (grid hashing) and count directed-edge defects: in a closed,
    consistently oriented surface every directed edge appears once and its reverse once."""
    v = np.asarray(soup, dtype=np.float64).reshape(-1, 3)
    # exact duplicates first, then clusters of distinct points closer than tol (union-find over
    # the pairs a k-d tree reports; rounding to a grid would split pairs that straddle a grid line)
    uniq, idx = np.unique(v, axis=0, return_inverse=True)
    idx = idx.reshape(-1)
    if tol > 0 and len(uniq) > 1:
        from scipy.spatial import cKDTree
        parent = np.arange(len(uniq))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i
        for a, b in cKDTree(uniq).query_pairs(tol):
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        roots = np.array([find(i) for i in range(len(uniq))])
        idx = roots[idx]
    tri = idx.reshape(-1, 3)
    tri = tri[(tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])]  # drop slivers merged away
    edges = {}
    for a, b, c in tri:
        for e in ((a, b), (b, c), (c, a)):
            edges[e] = edges.get(e, 0) + 1
    bad = 0
    for (a, b), n in edges.items():
        if n != 1 or edges.get((b, a), 0) != 1:
            bad += 1
    return {"triangles": len(tri), "vertices": len(np.unique(idx)), "bad_edges": bad}


def signed_volume(soup):
    s = np.asarray(soup, dtype=np.float64).reshape(-1, 3, 3)
    return float(np.einsum("ij,ij->i", s[:, 0], np.cross(s[:, 1], s[:, 2])).sum() / 6.0)
