"""Bitmap of a 2-D shape — same entry point as the reference's
/root/reference/codecad/rendering/bitmap.py:12-30 `render(obj, size)` -> uint8 [height][width][3].
Kernel: `cc_bitmap` (csrc/cc_kernels.cu, rendering/bitmap.cl:1-18)."""
import ctypes

import numpy as np

from .. import _lib
from ..cl_util.buffer import _Pinned
from ..geometry import Vector
from ..nodes import make_program_buffer


def render(obj, size):
    if hasattr(obj, "check_dimension"):
        obj.check_dimension(required=2)
    else:
        assert obj.dimension() == 2, "bitmap needs a 2D shape"
    box = obj.bounding_box().flattened()
    box_size = box.size()
    resolution = Vector(size[0], size[1], 1)  # the final 1 avoids a division by zero
    step_size = box_size.elementwise_div(resolution).max()
    origin = box.midpoint() - resolution * step_size / 2

    program = make_program_buffer(obj)
    w, h = int(size[0]), int(size[1])
    out = _Pinned(w * h * 3).array(np.uint8, (w, h, 3))  # pooled page-locked memory: the copy-out runs at PCIe speed
    L = _lib.lib()
    d_out = ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(out.nbytes, ctypes.byref(d_out)))
    try:
        o = (ctypes.c_float * 3)(*origin.as_float4().tolist()[:3])
        _lib.check(L.cc_bitmap(program.handle, o, ctypes.c_float(np.float32(step_size)), w, h, d_out, None))
        _lib.check(L.cc_memcpy_d2h_async(out.ctypes.data, d_out, out.nbytes, None))
        _lib.check(L.cc_synchronize())
    finally:
        L.cc_buffer_free(d_out)
    return out.transpose((1, 0, 2))
