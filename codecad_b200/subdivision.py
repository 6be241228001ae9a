"""Adaptive block subdivision — same signatures and results as the reference's
/root/reference/codecad/subdivision.py (calculate_block_sizes :116-166, subdivision
:169-253); the per-block `_Helper` launch/readback loop (:14-113) is replaced by one
device-resident pass per level inside libcodecad_b200 (cc_subdivide).
"""
import collections.abc
import ctypes
import itertools
import math

import numpy as np

from . import _lib
from .geometry import BoundingBox, Vector, as_vector
from .nodes import make_program_buffer


def calculate_block_sizes(box, dimension, resolution, grid_size, overlap, level_size_multiplier=1):
    """Level plan of the hierarchy, top level first: [(cell size in leaf cells, Vector grid dims), ...].
    Same contract as the reference's planner (subdivision.py:116-166; its 80 property cases are
    tests/test_host_logic.py): the leaf level has cell size 1; each level above multiplies the
    cell size by grid_size, except that with `overlap` a leaf block advances by grid_size - 1
    cells (neighbouring leaf blocks share a layer of samples); every level below the top is a
    full grid (flat in z for 2-D shapes); the top level is only as large as the box needs,
    rounded up to level_size_multiplier."""
    if grid_size % level_size_multiplier:
        raise ValueError("Grid size must be divisible by level_size_multiplier")
    assert dimension in (2, 3)
    lo, hi = as_vector(box[0]), as_vector(box[1])
    full = (grid_size, grid_size, grid_size)
    if dimension == 2:
        lo, hi = lo.flattened(), hi.flattened()
        full = (grid_size, grid_size, 1)
    samples = [math.ceil(extent / resolution) for extent in (hi - lo)]  # leaf cells per axis
    # cell sizes from the leaf upwards, until one block of the level spans the longest axis
    cells, factor = [1], (grid_size - 1 if overlap else grid_size)
    while cells[-1] * factor < max(samples):
        cells.append(cells[-1] * factor)
        factor = grid_size
    top = cells.pop()
    shared_layer = 1 if (overlap and not cells) else 0  # a single-level plan keeps the extra sample layer
    top_dims = []
    for n, limit in zip(samples, full):
        want = math.ceil(n / top) + shared_layer
        want = -(-want // level_size_multiplier) * level_size_multiplier
        top_dims.append(min(max(want, 1), limit))
    return [(top, Vector(*top_dims))] + [(c, Vector(*full)) for c in reversed(cells)]


try:  # optional C helper (csrc/cc_pylist.c) that builds the list of block tuples
    from . import _cc_pylist as _pylist
except ImportError:
    _pylist = None


def _levels(block_sizes):
    arr = (_lib.Level * len(block_sizes))()
    for i, (cell, dims) in enumerate(block_sizes):
        arr[i].cell_size = int(cell)
        arr[i].nx, arr[i].ny, arr[i].nz = int(dims[0]), int(dims[1]), int(dims[2])
    return arr


def subdivide_int_corners(program, origin, resolution, block_sizes, dimension, rank=0, world=1):
    """int64 array [n][3]: int corners (resolution units) of this rank's leaf blocks, in
    deterministic breadth-first order."""
    out = ctypes.POINTER(ctypes.c_int64)()
    count = ctypes.c_uint64()
    org = (ctypes.c_double * 3)(float(origin[0]), float(origin[1]), float(origin[2]))
    _lib.check(_lib.lib().cc_subdivide(program.handle, org, float(resolution), _levels(block_sizes),
                                       len(block_sizes), int(dimension), int(rank), int(world),
                                       ctypes.byref(out), ctypes.byref(count)))
    n = int(count.value)
    if n == 0:
        return np.zeros((0, 3), np.int64)
    try:
        return np.ctypeslib.as_array(out, shape=(n, 3)).copy()
    finally:
        _lib.lib().cc_free(out)


def sort_leaf_corners(corners, block_sizes):
    """Put int corners gathered from several ranks into the order one GPU lists them (level by
    level: parent block order, then INDEX3 cell order); returns a new int64 [n][3] array."""
    c = np.ascontiguousarray(corners, dtype=np.int64).reshape(-1, 3).copy()
    if len(c) > 1 and len(block_sizes) >= 2:
        _lib.check(_lib.load().cc_sort_leaf_corners(c.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(c),
                                                    _levels(block_sizes), len(block_sizes)))
    return c


class LeafBlocks(collections.abc.Sequence):
    """The third element of subdivision()'s result: a read-only sequence of
    (grid_dims, corner Vector(float64), step, int_corner Vector(int), int_step) tuples, exactly the
    reference's list (subdivision.py:97-111), built on demand from two arrays.  12 000 leaf blocks
    are 36 000 Python objects; callers that hand all blocks to the device in one go
    (rendering.mesh, rendering.polygon2d) read `.corners` / `.int_corners` and never create them."""

    def __init__(self, dims, corners, step, int_corners, int_step):
        self.dims, self.step, self.int_step = dims, step, int_step
        self.corners = corners          # float64 [n][3]: int_pos * resolution + origin
        self.int_corners = int_corners  # int64 [n][3]
        self._items = None

    def _materialise(self):
        if self._items is None and _pylist is not None:
            self._items = _pylist.leaf_blocks(Vector, self.dims, np.ascontiguousarray(self.corners, dtype=np.float64),
                                              self.step, np.ascontiguousarray(self.int_corners, dtype=np.int64),
                                              self.int_step)
        if self._items is None:
            new, vec = tuple.__new__, itertools.repeat(Vector)
            self._items = list(zip(itertools.repeat(self.dims), map(new, vec, self.corners.tolist()),
                                   itertools.repeat(self.step), map(new, vec, self.int_corners.tolist()),
                                   itertools.repeat(self.int_step)))
        return self._items

    def __len__(self):
        return len(self.int_corners)

    def __getitem__(self, i):
        if self._items is None and isinstance(i, (int, np.integer)):
            i = int(i)
            if not -len(self) <= i < len(self):
                raise IndexError("block index out of range")
            return (self.dims, Vector._make(self.corners[i].tolist()), self.step,
                    Vector._make(self.int_corners[i].tolist()), self.int_step)
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return "LeafBlocks(%d blocks of %s)" % (len(self), tuple(self.dims))


def block_corners(blocks):
    """float64 [n][3] box corners of a subdivision() block list (without building the tuples)."""
    if isinstance(blocks, LeafBlocks):
        return blocks.corners
    return np.array([[b[1][0], b[1][1], b[1][2]] for b in blocks], dtype=np.float64).reshape(-1, 3)


def subdivision(shape, resolution, overlap_edge_samples=True, grid_size=None, rank=0, world=1):
    """subdivision.py:169-253.  Returns (program_buffer, max_grid_dims,
    [(grid_dims, corner, step, int_corner, int_step), ...]) for the leaf blocks.

    `rank` / `world` (not in the reference) select the share of one rank when the first
    refined level is dealt round-robin over several GPUs; the union over ranks is the
    reference's block set."""
    if grid_size is None:
        grid_size = 128

    assert resolution > 0, "Non-positive resolution makes no sense"
    assert grid_size > 1, "Grid needs to be at least 2x2x2"
    assert grid_size <= 256, "Grid size > 256 would cause overflows in returned index list."

    program_buffer = make_program_buffer(shape)

    bb = shape.bounding_box()
    box = BoundingBox(as_vector(bb[0]), as_vector(bb[1])).expanded_additive(resolution / 2)
    dimension = shape.dimension()
    if dimension == 2:
        box = box.flattened()

    block_sizes = calculate_block_sizes(box, dimension, resolution, grid_size, overlap_edge_samples)

    if len(block_sizes) == 1:
        blocks = [(block_sizes[0][1], box.a, resolution, Vector(0, 0, 0), 1)] if rank == 0 else []
        return program_buffer, block_sizes[0][1], blocks

    corners = subdivide_int_corners(program_buffer, box.a, resolution, block_sizes, dimension, rank, world)
    leaf_step_int, leaf_dims = block_sizes[-1]
    leaf_step = leaf_step_int * resolution
    # subdivision.py:101  pos = int_pos * resolution + origin   (float64, same two IEEE operations
    # per axis as the reference's Vector arithmetic, vectorised; tuples built with _make)
    pos = corners * float(resolution) + np.array([box.a.x, box.a.y, box.a.z], dtype=np.float64)
    return program_buffer, leaf_dims, LeafBlocks(leaf_dims, pos, leaf_step, corners, leaf_step_int)
