"""Slice viewer — /root/reference/codecad/rendering/matplotlib_slice.py:14-96.

`slice_values(obj)` is the device half (grid set-up :19-52, kernel matplotlib_slice.cl:1-20): the
float32 array [height][width][3] of (distance, gradient x, gradient y) over the z = 0 plane of the
bounding box, plus the grid's corner and resolution.  `render_slice(obj)` plots it like the reference
and needs matplotlib (not part of this image; imported on use)."""
import ctypes
import math

import numpy as np

from .. import _lib
from ..geometry import BoundingBox, Vector
from ..nodes import make_program_buffer


def slice_values(obj):
    resolution = obj.feature_size() / 2
    bb = obj.bounding_box()
    grow = (bb.b - bb.a) * 0.1                                   # BoundingBox.expanded(0.1), util/geometry.py:150-153
    box = BoundingBox(bb.a - grow, bb.b + grow).flattened()
    box_size = box.size()
    grid_dimensions = [math.ceil(s / resolution) + 1 for s in box_size]
    new_box_size = Vector(*(resolution * (k - 1) for k in grid_dimensions)).flattened()
    grid_dimensions[2] = 3
    corner = box.midpoint() - new_box_size / 2

    program = make_program_buffer(obj)
    values = np.empty((grid_dimensions[1], grid_dimensions[0], grid_dimensions[2]), dtype=np.float32)
    L = _lib.lib()
    d_out = ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(values.nbytes, ctypes.byref(d_out)))
    try:
        _lib.check(L.cc_matplotlib_slice(program.handle, _lib.f3(corner.as_float4()), float(np.float32(resolution)),
                                         grid_dimensions[0], grid_dimensions[1], d_out, None))
        _lib.check(L.cc_memcpy_d2h_async(values.ctypes.data, d_out, values.nbytes, None))
        _lib.check(L.cc_synchronize())
    finally:
        L.cc_buffer_free(d_out)
    return values, corner, resolution, new_box_size


def render_slice(obj, _filename=None):  # second argument: interface compatibility with other renderers
    import matplotlib
    import matplotlib.pyplot as plt
    values, corner, resolution, new_box_size = slice_values(obj)
    distances = values[:, :, 0]
    distance_range = np.max(np.abs(distances))
    common_args = {
        "norm": matplotlib.colors.SymLogNorm(0.1, vmin=-distance_range, vmax=distance_range),
        "origin": "lower", "aspect": "equal",
        "extent": (corner.x, corner.x + (values.shape[1] - 1) * resolution,
                   corner.y, corner.y + (values.shape[0] - 1) * resolution),
    }
    plt.imshow(distances, cmap=plt.get_cmap("RdBu"), interpolation="none", **common_args)
    plt.colorbar()
    plt.contour(distances, colors="black", **common_args)
    thin = 10
    plt.quiver(np.arange(0, values.shape[1], thin) * resolution + corner.x,
               np.arange(0, values.shape[0], thin) * resolution + corner.y,
               values[::thin, ::thin, 1], values[::thin, ::thin, 2],
               color="white", angles="xy", scale_units="xy", scale=200 / new_box_size.max())
    plt.show()
