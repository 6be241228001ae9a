"""Boundary polygons of a 2-D shape — same entry point as the reference's
/root/reference/codecad/rendering/polygon2d.py:36-173 `polygon(obj, subdivision_grid_size=None)`:
a generator of closed vertex chains [(x, y), ...].

The reference walks the leaf boxes of `subdivision(obj, feature_size / 2)` one at a time: a
grid_eval launch, a process_polygon launch and five blocking reads per box (polygon2d.py:88-117).
Here `cc_polygon_blocks` evaluates and processes every box in two launches and one round of
copies; `cc_polygon_assemble` (native host code, no device work) then follows the links exactly as
the reference's Python does:

  * closed chains inside a box come out in increasing order of their first triangle
    (polygon2d.py:165-170) — for a shape that fits one box (the default grid of 128 and
    resolution = feature_size / 2 make that the common case) the output is the reference's;
  * chains that leave a box are pieces keyed by (box, side + row) at both ends and are joined to
    the pieces of the neighbouring boxes, in whatever order the boxes arrive, until they close.
    (The reference's bookkeeping for this case, polygon2d.py:131-163, leaves stale entries behind
    as soon as a chain crosses more than one box boundary and then trips its own assertion; the
    piece table of the native code implements what it is meant to do.)
"""
import ctypes

import numpy as np

from .. import _lib
from ..cl_util.buffer import _Pinned
from ..subdivision import LeafBlocks, block_corners, subdivision


def polygon_blocks(program_buffer, grid, corners, resolution):
    """Device half: grid = (gx, gy) samples per box, corners float64 [n][3].
    -> (vertices float32 [n][cells][2], links uint32 [n][cells], starts uint32 [n][gx+gy-2], counts uint32 [n])
    with cells = 2*(gx-1)*(gy-1) in INDEX3 order t + 2*(y + (gy-1)*x)."""
    gx, gy = int(grid[0]), int(grid[1])
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 3)
    n = len(corners)
    cells = 2 * (gx - 1) * (gy - 1)
    if n == 0:
        return (np.zeros((0, cells, 2), np.float32), np.zeros((0, cells), np.uint32),
                np.zeros((0, gx + gy - 2), np.uint32), np.zeros((0,), np.uint32))
    # page-locked (pooled) result arrays: every element is written by the copies
    vertices = _Pinned(n * cells * 8).array(np.float32, (n, cells, 2))
    links = _Pinned(n * cells * 4).array(np.uint32, (n, cells))
    starts = _Pinned(n * (gx + gy - 2) * 4).array(np.uint32, (n, gx + gy - 2))
    counts = _Pinned(n * 4).array(np.uint32, (n,))
    if n:
        _lib.check(_lib.lib().cc_polygon_blocks(
            program_buffer.handle, corners.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(resolution),
            gx, gy, n, vertices.ctypes.data, links.ctypes.data, starts.ctypes.data, counts.ctypes.data))
    return vertices, links, starts, counts


ERR_OPEN_OUTLINE = -6  # CC_ERR_OPEN_OUTLINE (include/codecad_b200.h)


def assemble(vertices, links, starts, counts, int_corners, int_step):
    """Host half (native, no device work): dense per-box arrays of polygon_blocks() -> list of
    closed chains [(x, y), ...] in the order they close.  `links` is consumed (visit marks)."""
    n, cells = links.shape
    int_corners = np.ascontiguousarray(int_corners, dtype=np.int64).reshape(-1, 3)
    assert len(int_corners) == n and vertices.shape == (n, cells, 2) and starts.shape[0] == n
    vertices = np.ascontiguousarray(vertices, dtype=np.float32)
    links = np.ascontiguousarray(links, dtype=np.uint32)
    starts = np.ascontiguousarray(starts, dtype=np.uint32)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    out_v = ctypes.POINTER(ctypes.c_float)()
    out_o = ctypes.POINTER(ctypes.c_uint64)()
    n_chains = ctypes.c_uint64()
    L = _lib.load()
    rc = L.cc_polygon_assemble(vertices.ctypes.data, links.ctypes.data, starts.ctypes.data, counts.ctypes.data,
                               int_corners.ctypes.data, int(int_step), cells, starts.shape[1], n,
                               ctypes.byref(out_v), ctypes.byref(out_o), ctypes.byref(n_chains))
    if rc == ERR_OPEN_OUTLINE:
        # the reference ends with `assert len(open_chain_beginnings) == 0` (polygon2d.py:172-173)
        raise AssertionError("an outline left the subdivided region")
    _lib.check(rc)
    try:
        k = int(n_chains.value)
        offsets = np.ctypeslib.as_array(out_o, shape=(k + 1,)).tolist()
        total = offsets[-1]
        flat = np.ctypeslib.as_array(out_v, shape=(max(total, 1), 2))[:total].tolist()
        return [list(map(tuple, flat[offsets[i]:offsets[i + 1]])) for i in range(k)]
    finally:
        L.cc_free(out_v)
        L.cc_free(out_o)


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
    if isinstance(boxes, LeafBlocks):
        int_corners, int_step = boxes.int_corners, boxes.int_step * (boxes.dims[0] - 1)
    else:
        int_corners = np.array([[b[3][0], b[3][1], b[3][2]] for b in boxes], dtype=np.int64)
        int_step = boxes[0][4] * (boxes[0][0][0] - 1)  # boxes share their border samples
    yield from assemble(vertices, links, starts, counts, int_corners, int_step)
