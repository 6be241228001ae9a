"""Host side of the reference's image renderers around the oracle's kernels — TEST
INFRASTRUCTURE ONLY.  Restates /root/reference/codecad/rendering/ray_caster.py:30-115
(camera set-up, kernel arguments), bitmap.py:12-30 and image.py:7-16, so that the oracle's
evaluate() can be compared with the reference's golden PNGs (tests/golden/reference_renders,
copied from the reference's tests/baseline: real outputs of its OpenCL path)."""
import ctypes
import math

import numpy as np

from . import lib, _f32, _fp, _u8p


def _v(a):
    return np.asarray(a, dtype=np.float64)


def _norm(a):
    return a / math.sqrt(float(a @ a))


def _zero_if_inf(x):
    return 0 if math.isinf(x) else x


def camera_params(box_a, box_b, size, view_angle=None):
    """ray_caster.py:92-115."""
    a, b = _v(box_a), _v(box_b)
    box_size = b - a
    diagonal = math.hypot(*size)
    focal = diagonal if view_angle is None else diagonal / (2 * math.tan(math.radians(view_angle) / 2))
    distance = focal * max(_zero_if_inf(box_size[0]) / size[0], _zero_if_inf(box_size[2]) / size[1])
    if distance == 0:
        distance = 1
    distance *= 1.2
    origin = (a + b) / 2 - np.array([0, distance + _zero_if_inf(box_size[1]) / 2, 0])
    return origin, np.array([0.0, 1.0, 0.0]), np.array([0.0, 0.0, 1.0]), focal


def ray_cast_args(box_a, box_b, size, view_angle=None):
    """The kernel arguments ray_caster.py:30-71 computes for a picture of the box: origin, forward (scaled
    by the focal length), up, right, pixel_tolerance, box_radius, min_distance, max_distance, floor_z."""
    a, b = _v(box_a), _v(box_b)
    origin, direction, up, focal = camera_params(a, b, size, view_angle)
    forward = _norm(direction)
    up = up - forward * float(up @ forward)
    up = _norm(up)
    right = np.cross(forward, up)
    forward = forward * focal
    mid = (a + b) / 2
    origin_to_mid = math.sqrt(float((origin - mid) @ (origin - mid)))
    box_radius = math.sqrt(float((b - a) @ (b - a))) / 2
    return (origin, forward, up, right, 0.5 / focal, box_radius, max(0, origin_to_mid - box_radius),
            origin_to_mid + box_radius, a[2] - (b[2] - a[2]) / 20)


def bitmap_args(box_a, box_b, size):
    """origin and step of bitmap.py:12-30."""
    a, b = _v(box_a).copy(), _v(box_b).copy()
    a[2] = b[2] = 0
    resolution = np.array([size[0], size[1], 1.0])
    step = float(((b - a) / resolution).max())
    return (a + b) / 2 - resolution * step / 2, step


def ray_cast(words, box_a, box_b, size, view_angle=None, options=0):
    """ray_caster.py:30-89 -> uint8 [h][w][3]."""
    a, b = _v(box_a), _v(box_b)
    origin, direction, up, focal = camera_params(a, b, size, view_angle)
    forward = _norm(direction)
    up = up - forward * float(up @ forward)
    up = _norm(up)
    right = np.cross(forward, up)
    forward = forward * focal
    pixel_tolerance = 0.5 / focal
    mid = (a + b) / 2
    origin_to_mid = math.sqrt(float((origin - mid) @ (origin - mid)))
    box_radius = math.sqrt(float((b - a) @ (b - a))) / 2
    min_distance = max(0, origin_to_mid - box_radius)
    max_distance = origin_to_mid + box_radius
    floor_z = a[2] - (b[2] - a[2]) / 20
    w = _f32(words)
    out = np.zeros((size[0], size[1], 3), np.uint8)
    f = lambda v: _f32(v).ctypes.data_as(_fp)
    rc = lib().oracle_ray_caster(w.ctypes.data_as(_fp), len(w), f(origin), f(forward), f(up), f(right),
                                 ctypes.c_float(pixel_tolerance), ctypes.c_float(box_radius),
                                 ctypes.c_float(min_distance), ctypes.c_float(max_distance), ctypes.c_float(floor_z),
                                 ctypes.c_uint(int(options)), size[0], size[1], out.ctypes.data_as(_u8p))
    assert rc == 0
    return out.transpose((1, 0, 2))


def bitmap(words, box_a, box_b, size):
    """bitmap.py:12-30 -> uint8 [h][w][3]."""
    a, b = _v(box_a).copy(), _v(box_b).copy()
    a[2] = b[2] = 0
    box_size = b - a
    resolution = np.array([size[0], size[1], 1.0])
    step = float((box_size / resolution).max())
    origin = (a + b) / 2 - resolution * step / 2
    w = _f32(words)
    out = np.zeros((size[0], size[1], 3), np.uint8)
    rc = lib().oracle_bitmap(w.ctypes.data_as(_fp), len(w), _f32(origin).ctypes.data_as(_fp), ctypes.c_float(step),
                             size[0], size[1], out.ctypes.data_as(_u8p))
    assert rc == 0
    return out.transpose((1, 0, 2))


def render(words, dimension, box_a, box_b, size=(1024, 768)):
    """image.py:7-16."""
    if dimension == 2:
        return bitmap(words, box_a, box_b, size)
    return ray_cast(words, box_a, box_b, size)
