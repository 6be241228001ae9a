"""Boundary polygons of a 2-D shape — same entry point as the reference's
/root/reference/codecad/rendering/polygon2d.py:36-173 `polygon(obj, subdivision_grid_size=None)`:
a generator of closed vertex chains [(x, y), ...].

The reference walks the leaf boxes of `subdivision(obj, feature_size / 2)` one at a time: a
grid_eval launch, a process_polygon launch and five blocking reads per box (polygon2d.py:88-117).
Here `cc_polygon_blocks` evaluates and processes every box in two launches and one round of
copies; the host then follows the links exactly as the reference does:

  * closed chains inside a box come out in increasing order of their first triangle
    (polygon2d.py:165-170) — for a shape that fits one box (the default grid of 128 and
    resolution = feature_size / 2 make that the common case) the output is the reference's;
  * chains that leave a box are pieces keyed by (box, side + row) at both ends and are joined to
    the pieces of the neighbouring boxes, in whatever order the boxes arrive, until they close.
    (The reference's bookkeeping for this case, polygon2d.py:131-163, leaves stale entries behind
    as soon as a chain crosses more than one box boundary and then trips its own assertion; the
    piece table below implements what it is meant to do.)
"""
import ctypes

import numpy as np

from .. import _lib
from ..geometry import Vector
from ..subdivision import LeafBlocks, block_corners, subdivision

_LINK_OVERFLOW_MASK = 0xFFF00000   # polygon2d.cl:5-35: flags (bits 29-31) + row / column (bits 20-28)
_INDEX_MASK = 0x000FFFFF


def polygon_blocks(program_buffer, grid, corners, resolution):
    """Device half: grid = (gx, gy) samples per box, corners float64 [n][3].
    -> (vertices float32 [n][cells][2], links uint32 [n][cells], starts uint32 [n][gx+gy-2], counts uint32 [n])
    with cells = 2*(gx-1)*(gy-1) in INDEX3 order t + 2*(y + (gy-1)*x)."""
    gx, gy = int(grid[0]), int(grid[1])
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 3)
    n = len(corners)
    cells = 2 * (gx - 1) * (gy - 1)
    vertices = np.zeros((n, cells, 2), np.float32)
    links = np.full((n, cells), 0xFFFFFFFF, np.uint32)
    starts = np.zeros((n, gx + gy - 2), np.uint32)
    counts = np.zeros((n,), np.uint32)
    if n:
        _lib.check(_lib.lib().cc_polygon_blocks(
            program_buffer.handle, corners.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(resolution),
            gx, gy, n, vertices.ctypes.data, links.ctypes.data, starts.ctypes.data, counts.ctypes.data))
    return vertices, links, starts, counts


def _follow(links, vertices, index, chain):
    """Append the vertices from `index` on until a link leaves the box or reaches a visited
    triangle; visited links are overwritten with the mask.  Returns the terminating link's flags."""
    while not index & _LINK_OVERFLOW_MASK:
        chain.append(vertices[index])
        nxt = links[index]
        links[index] = _LINK_OVERFLOW_MASK
        index = nxt
    return index & _LINK_OVERFLOW_MASK


def _neighbour_step(spec):
    """Which neighbouring box a link with flags `spec` leads into (polygon2d.py:26-31)."""
    d = -1 if spec & 0x20000000 else 1
    return (0, d) if spec & 0x40000000 else (d, 0)


class _Piece:
    __slots__ = ("chain", "begin", "end")

    def __init__(self, chain, begin, end):
        self.chain, self.begin, self.end = chain, begin, end


class _OpenChains:
    """Pieces of outlines that cross box boundaries, joinable at either end."""

    def __init__(self):
        self.by_begin = {}
        self.by_end = {}

    def add(self, chain, begin, end):
        """Returns a closed chain if this piece completed one, else None."""
        prev = self.by_end.pop(begin, None)
        if prev is None:
            piece = _Piece(chain, begin, end)
            self.by_begin[begin] = piece
        else:  # continues a chain that ended on this box's border
            prev.chain.extend(chain)
            prev.end = end
            piece = prev
        nxt = self.by_begin.pop(end, None)
        if nxt is None:
            self.by_end[end] = piece
            return None
        if nxt is piece:
            return piece.chain
        piece.chain.extend(nxt.chain)
        piece.end = nxt.end
        self.by_end[nxt.end] = piece
        return None

    def __len__(self):
        return len(self.by_begin) + len(self.by_end)


def polygon(obj, subdivision_grid_size=None):
    """ Generate polygons representing the boundaries of a 2D shape. """
    if hasattr(obj, "check_dimension"):
        obj.check_dimension(required=2)
    else:
        assert obj.dimension() == 2, "polygon needs a 2D shape"

    program_buffer, grid_size, boxes = subdivision(obj, obj.feature_size() / 2, grid_size=subdivision_grid_size)

    assert grid_size[0] < 512, "Larger grid size would overflow the index encoding"
    assert grid_size[2] == 1
    if not len(boxes):
        return
    if len(boxes) > 1:
        for box_size in ([boxes.dims] if isinstance(boxes, LeafBlocks) else [b[0] for b in boxes]):
            assert box_size[0] == box_size[1]

    vertices, links, starts, counts = polygon_blocks(program_buffer, (grid_size[0], grid_size[1]),
                                                     block_corners(boxes), boxes[0][2])
    open_chains = _OpenChains()
    # only the triangles the outline crosses matter to the host: pull them out of the dense
    # per-box arrays in one pass (links of untouched triangles are 0xFFFFFFFF)
    hit_box, hit_index = np.nonzero(links != 0xFFFFFFFF)
    first = np.searchsorted(hit_box, np.arange(len(boxes) + 1))
    hit_links = links[hit_box, hit_index].tolist()
    hit_vertices = list(map(tuple, vertices[hit_box, hit_index].tolist()))
    hit_index = hit_index.tolist()
    counts = counts.tolist()
    for b in range(len(boxes)):
        lo, hi = int(first[b]), int(first[b + 1])
        if lo == hi:
            continue
        box_size, _, _, int_corner, int_resolution = boxes[b]
        surface = hit_index[lo:hi]
        box_links = dict(zip(surface, hit_links[lo:hi]))
        box_vertices = dict(zip(surface, hit_vertices[lo:hi]))
        assert counts[b] <= starts.shape[1]
        if counts[b]:
            int_step = int_resolution * (box_size[0] - 1)  # boxes share their border samples
            for start in starts[b, :counts[b]].tolist():
                begin = (int_corner[0], int_corner[1], start & _LINK_OVERFLOW_MASK)
                chain = []
                spec = _follow(box_links, box_vertices, start & _INDEX_MASK, chain)
                dx, dy = _neighbour_step(spec)
                end = (int_corner[0] + dx * int_step, int_corner[1] + dy * int_step, spec)
                closed = open_chains.add(chain, begin, end)
                if closed is not None:
                    yield closed
        for index in surface:  # what is left belongs to chains closed inside the box
            if box_links[index] & _LINK_OVERFLOW_MASK:
                continue
            chain = []
            _follow(box_links, box_vertices, index, chain)
            yield chain
    assert len(open_chains) == 0, "an outline left the subdivided region"
