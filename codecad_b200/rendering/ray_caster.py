"""Ray-cast picture of a 3-D shape — same entry points as the reference's
/root/reference/codecad/rendering/ray_caster.py:12-115: `RenderOptions`, `render(obj, origin,
direction, up, focal_length, size, options)` -> uint8 [height][width][3], `get_camera_params(box,
size, view_angle)`.

The host arithmetic (camera basis, pixel tolerance, clipping distances, floor height) is the
reference's, in float64, rounded to fp32 where the reference hands it to the kernel
(`as_float4()`, `numpy.float32`).  The kernel is `cc_ray_caster` (csrc/cc_kernels.cu): the same
SDF interpreter as grid_eval under a warp-synchronous ray-marching state machine.
"""
import ctypes
import enum
import math

import numpy as np

from .. import _lib
from ..cl_util.buffer import _Pinned
from ..geometry import Vector, as_vector
from ..nodes import make_program_buffer


class RenderOptions(enum.IntFlag):
    """ray_caster.py:12-15 (py-flags there; the integer values are what the kernel sees)."""
    no_flags = 0
    false_color = 1
    zebra = 2


def _zero_if_inf(x):
    return 0 if math.isinf(x) else x


def _check_3d(obj):
    if hasattr(obj, "check_dimension"):
        obj.check_dimension(required=3)
    else:
        assert obj.dimension() == 3, "the ray caster needs a 3D shape"


def _f3(v):
    return (ctypes.c_float * 3)(*np.asarray(v.as_float4().tolist()[:3], dtype=np.float32))


def kernel_arguments(box, origin, direction, up, focal_length):
    """ray_caster.py:36-50: what the reference computes between the call and the kernel launch."""
    origin, direction, up = as_vector(origin), as_vector(direction), as_vector(up)
    forward = direction.normalized()
    up = up - forward * up.dot(forward)
    up = up.normalized()
    right = forward.cross(up)
    forward = forward * focal_length
    pixel_tolerance = 0.5 / focal_length  # tangent of half a pixel's angle
    origin_to_midpoint = abs(origin - box.midpoint())
    box_radius = abs(box.size()) / 2
    min_distance = max(0, origin_to_midpoint - box_radius)
    max_distance = origin_to_midpoint + box_radius
    floor_z = box.a.z - box.size().z / 20
    return origin, forward, up, right, pixel_tolerance, box_radius, min_distance, max_distance, floor_z


def render(obj, origin, direction, up, focal_length, size, options=RenderOptions.no_flags, stats=None):
    """ray_caster.py:30-89.  `stats` (optional dict) receives 'evaluations' and 'ms'."""
    box = obj.bounding_box()
    _check_3d(obj)
    origin, forward, up, right, tol, radius, dmin, dmax, floor_z = kernel_arguments(box, origin, direction, up,
                                                                                    focal_length)
    program = make_program_buffer(obj)
    w, h = int(size[0]), int(size[1])
    out = _Pinned(w * h * 3).array(np.uint8, (w, h, 3))  # pooled page-locked memory: the copy-out runs at PCIe speed
    L = _lib.lib()
    d_out = ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(out.nbytes, ctypes.byref(d_out)))
    try:
        evals = ctypes.c_uint64()
        e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(L.cc_event_record(ctypes.byref(e0)))
        _lib.check(L.cc_ray_caster(program.handle, _f3(origin), _f3(forward), _f3(up), _f3(right),
                                   ctypes.c_float(tol), ctypes.c_float(radius), ctypes.c_float(dmin),
                                   ctypes.c_float(dmax), ctypes.c_float(floor_z), int(options), w, h, d_out,
                                   ctypes.byref(evals) if stats is not None else None, ctypes.byref(e1)))
        _lib.check(L.cc_memcpy_d2h_async(out.ctypes.data, d_out, out.nbytes, None))
        _lib.check(L.cc_synchronize())
        if stats is not None:
            ms = ctypes.c_float()
            _lib.check(L.cc_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
            stats["evaluations"] = int(evals.value)
            stats["ms"] = float(ms.value)
        L.cc_event_destroy(e0)
        L.cc_event_destroy(e1)
    finally:
        L.cc_buffer_free(d_out)
    return out.transpose((1, 0, 2))


def get_camera_params(box, size, view_angle):
    """ray_caster.py:92-115."""
    box_size = box.size()
    size_diagonal = math.hypot(*size)
    if view_angle is None:
        focal_length = size_diagonal  # normal lens by default
    else:
        focal_length = size_diagonal / (2 * math.tan(math.radians(view_angle) / 2))
    distance = focal_length * max(_zero_if_inf(box_size.x) / size[0], _zero_if_inf(box_size.z) / size[1])
    if distance == 0:
        distance = 1
    distance *= 1.2  # 20 % margin around the object
    origin = box.midpoint() - Vector(0, distance + _zero_if_inf(box_size.y) / 2, 0)
    return (origin, Vector(0, 1, 0), Vector(0, 0, 1), focal_length)
