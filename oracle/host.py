"""TEST INFRASTRUCTURE — restatement of the reference's *host* algorithms on top of
the C oracle's per-block kernels.  float64 Python arithmetic, same operation order
as the reference, so block corners and integrals can be compared exactly.

  calculate_block_sizes  /root/reference/codecad/subdivision.py:116-166
  subdivision            /root/reference/codecad/subdivision.py:48-113,169-253
  mass_properties        /root/reference/codecad/mass_properties.py:30-229
  KahanSummation         /root/reference/codecad/util/math.py:4-21
"""
import math

import numpy as np

from . import mass_properties_step, subdivision_step


def _round_up_to(x, y):
    return ((x + y - 1) // y) * y


def _clamp(v, lo, hi):
    return max(lo, min(v, hi))


def calculate_block_sizes(box_a, box_b, dimension, resolution, grid_size, overlap, level_size_multiplier=1):
    """-> top..leaf list of (cell_size_in_leaf_units, (nx, ny, nz))."""
    if grid_size % level_size_multiplier != 0:
        raise ValueError("Grid size must be divisible by level_size_multiplier")
    if dimension == 2:
        level_size = (grid_size, grid_size, 1)
        box_a = (box_a[0], box_a[1], 0)
        box_b = (box_b[0], box_b[1], 0)
    elif dimension == 3:
        level_size = (grid_size,) * 3
    else:
        raise AssertionError
    box_int_size = tuple(math.ceil((b - a) / resolution) for a, b in zip(box_a, box_b))
    box_max_int_size = max(box_int_size)
    cell_size = 1
    block_sizes = []
    while True:
        block_sizes.append((cell_size, level_size))
        overlap_delta = 1 if overlap and len(block_sizes) == 1 else 0
        next_cell_size = cell_size * (grid_size - overlap_delta)
        if next_cell_size >= box_max_int_size:
            break
        cell_size = next_cell_size
    block_sizes[-1] = (
        cell_size,
        tuple(
            _clamp(_round_up_to(math.ceil(x / cell_size) + overlap_delta, level_size_multiplier), 1, s)
            for x, s in zip(box_int_size, level_size)
        ),
    )
    block_sizes.reverse()
    return block_sizes


def subdivision(words, box_a, box_b, dimension, resolution, overlap_edge_samples=True, grid_size=128):
    """-> (max_dims, [(dims, corner(float64 x3), step, int_corner(int x3), int_step)])
    with leaf blocks in canonical order (depth-first, INDEX3 order inside a block)."""
    assert resolution > 0 and 1 < grid_size <= 256
    a = tuple(c - resolution / 2 for c in box_a)
    b = tuple(c + resolution / 2 for c in box_b)
    if dimension == 2:
        a = (a[0], a[1], 0)
        b = (b[0], b[1], 0)
    bs = calculate_block_sizes(a, b, dimension, resolution, grid_size, overlap_edge_samples)
    if len(bs) == 1:
        return bs[0][1], [(bs[0][1], a, resolution, (0, 0, 0), 1)]
    final = []

    def visit(int_corner, level):
        int_step, dims = bs[level]
        if dimension == 3:
            shifted = tuple(c + int_step / 2 for c in int_corner)
        else:
            shifted = (int_corner[0] + int_step / 2, int_corner[1] + int_step / 2, int_corner[2] + 0)
        box_step = int_step * resolution
        corner = tuple(s * resolution + o for s, o in zip(shifted, a))
        thr = box_step * math.sqrt(dimension) / 2
        hits = subdivision_step(words, np.array(corner, np.float64).astype(np.float32), box_step, thr, dims)
        nxt = level + 1
        for i, j, k, _ in hits.tolist():
            pos = (i * int_step + int_corner[0], j * int_step + int_corner[1], k * int_step + int_corner[2])
            if nxt == len(bs) - 1:
                final.append((bs[nxt][1], tuple(p * resolution + o for p, o in zip(pos, a)),
                              bs[nxt][0] * resolution, pos, bs[nxt][0]))
            else:
                visit(pos, nxt)

    visit((0, 0, 0), 0)
    return bs[-1][1], final


class _Kahan:
    def __init__(self):
        self.result = 0.0
        self.correction = 0.0

    def add(self, x):
        y = x - self.correction
        tmp = self.result + y
        self.correction = (tmp - self.result) - y
        self.result = tmp


def mass_properties(words, box_a, box_b, resolution, grid_size=64, return_stats=False):
    """-> (volume, centroid(3), inertia 3x3) following mass_properties.py:30-229."""
    assert resolution > 0 and grid_size > 1 and grid_size ** 5 <= 2 ** 32
    bs = calculate_block_sizes(box_a, box_b, 3, resolution, grid_size, False)
    bs = [(resolution * cs, dims) for cs, dims in bs]
    acc = [_Kahan() for _ in range(10)]  # one, x, y, z, xx, yy, zz, xy, xz, yz
    stats = {"launches": 0, "evaluations": 0, "levels": len(bs)}
    stack = [(tuple(float(c) for c in box_a), 0)]
    while stack:
        corner, level = stack.pop()
        s, dims = bs[level]
        shifted = tuple(c + s / 2 for c in corner)
        thr = s * math.sqrt(3) / 2 if level < len(bs) - 1 else 0
        sums, hits = mass_properties_step(
            words, np.array(shifted, np.float64).astype(np.float32), s, thr, dims
        )
        stats["launches"] += 1
        stats["evaluations"] += dims[0] * dims[1] * dims[2]
        sxx, sxy, sxz, sx, syy, syz, sy, szz, sz, n = (int(v) for v in sums)
        s2 = s * s
        s3 = s * s2
        bx, by, bz = shifted
        tx, ty, tz = s * sx, s * sy, s * sz
        txx, tyy, tzz = s2 * sxx, s2 * syy, s2 * szz
        txy, txz, tyz = s2 * sxy, s2 * sxz, s2 * syz
        acc[0].add(s3 * n)
        acc[1].add(s3 * (n * bx + tx))
        acc[2].add(s3 * (n * by + ty))
        acc[3].add(s3 * (n * bz + tz))
        acc[4].add(s3 * (n * (bx * bx + s2 / 12) + 2 * bx * tx + txx))
        acc[5].add(s3 * (n * (by * by + s2 / 12) + 2 * by * ty + tyy))
        acc[6].add(s3 * (n * (bz * bz + s2 / 12) + 2 * bz * tz + tzz))
        acc[7].add(s3 * (n * bx * by + bx * ty + by * tx + txy))
        acc[8].add(s3 * (n * bx * bz + bx * tz + bz * tx + txz))
        acc[9].add(s3 * (n * by * bz + by * tz + bz * ty + tyz))
        assert level + 1 < len(bs) or len(hits) == 0
        stack.extend(
            ((i * s + corner[0], j * s + corner[1], k * s + corner[2]), level + 1)
            for i, j, k, _ in hits.tolist()
        )
    res = finish_mass_properties([a.result for a in acc])
    return res + (stats,) if return_stats else res


def finish_mass_properties(integrals):
    """mass_properties.py:179-229: integrals (one,x,y,z,xx,yy,zz,xy,xz,yz) ->
    (volume, centroid, inertia tensor about the centroid)."""
    one, ix, iy, iz, ixx, iyy, izz, ixy, ixz, iyz = (float(v) for v in integrals)
    if one == 0:
        return 0, (0, 0, 0), np.zeros((3, 3))
    cx, cy, cz = ix / one, iy / one, iz / one
    sxx = ixx - 2 * cx * ix + cx * cx * one
    syy = iyy - 2 * cy * iy + cy * cy * one
    szz = izz - 2 * cz * iz + cz * cz * one
    sxy = ixy - cx * iy - cy * ix + cx * cy * one
    sxz = ixz - cx * iz - cz * ix + cx * cz * one
    syz = iyz - cy * iz - cz * iy + cy * cz * one
    Ixx, Iyy, Izz = syy + szz, sxx + szz, sxx + syy
    Ixy, Ixz, Iyz = -sxy, -sxz, -syz
    return one, (cx, cy, cz), np.array([[Ixx, Ixy, Ixz], [Ixy, Iyy, Iyz], [Ixz, Iyz, Izz]])


# ---- rendering/polygon2d.py:36-173 ---------------------------------------------------------------

_POLY_MASK = 0xFFF00000


def polygon(words, box_a, box_b, feature_size, grid_size=128):
    """Boundary polygons of a 2-D scene: subdivision(feature_size / 2) -> per box grid_eval +
    process_polygon (polygon2d.py:80-117) -> chains.  Closed chains inside a box are collected like
    polygon2d.py:165-170.  Pieces that cross box borders are gathered from ALL boxes first and then
    linked end-key -> begin-key (the keys of polygon2d.py:131,141-144); this is deliberately a
    different procedure from the product's online joining, and from the reference's, whose
    bookkeeping does not survive a chain that crosses two borders (see codecad_b200/rendering/
    polygon2d.py)."""
    from . import grid_eval, process_polygon
    max_dims, boxes = subdivision(words, box_a, box_b, 2, feature_size / 2, True, grid_size)
    assert max_dims[0] < 512 and max_dims[2] == 1
    closed, pieces = [], {}
    for dims, corner, step, int_corner, int_step in boxes:
        c32 = np.array(corner, np.float64).astype(np.float32)
        field = grid_eval(words, c32, step, (max_dims[0], max_dims[1], 1))[:, :, 0, :]
        vertices, links, starts = process_polygon(c32, step, field)
        links = links.tolist()
        vertices = [tuple(v) for v in vertices.tolist()]

        def collect(i, chain):
            while not i & _POLY_MASK:
                chain.append(vertices[i])
                last = i
                i = links[i]
                links[last] = _POLY_MASK
            return i & _POLY_MASK

        box_step = int_step * (dims[0] - 1)
        for s in starts.tolist():
            chain = []
            spec = collect(s & ~_POLY_MASK & 0xFFFFFFFF, chain)
            d = -1 if spec & 0x20000000 else 1
            step_xy = (0, d) if spec & 0x40000000 else (d, 0)
            begin = (int_corner[0], int_corner[1], s & _POLY_MASK)
            end = (int_corner[0] + step_xy[0] * box_step, int_corner[1] + step_xy[1] * box_step, spec)
            assert begin not in pieces
            pieces[begin] = (chain, end)
        for i in range(len(links)):
            if links[i] & _POLY_MASK:
                continue
            chain = []
            collect(i, chain)
            closed.append(chain)
    while pieces:
        begin = min(pieces)
        chain, end = pieces.pop(begin)
        chain = list(chain)
        while end != begin:
            nxt, end = pieces.pop(end)  # KeyError = an outline that never closes
            chain.extend(nxt)
        closed.append(chain)
    return closed
