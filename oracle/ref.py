"""TEST INFRASTRUCTURE — ctypes front end of oracle/_ref/libcodecad_ref.so, the
reference's own OpenCL device sources compiled for the host (oracle/build_ref.py).
Same call shapes as the oracle port in oracle/__init__.py."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libcodecad_ref.so")
_lib = None
_fp = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_SO)
        L.ref_grid_eval.argtypes = [_fp, _fp, ctypes.c_float] + [ctypes.c_int] * 3 + [_fp]
        L.ref_grid_eval_pymcubes.argtypes = [_fp, _fp, ctypes.c_float] + [ctypes.c_int] * 3 + [_fp]
        L.ref_subdivision_step.argtypes = [_fp, _fp, ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 3 + [_u32p, _u8p]
        L.ref_mass_properties.argtypes = [_fp, _fp, ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 3 + [_u32p, _u32p, _u8p]
        L.ref_num_threads.restype = ctypes.c_int
        if hasattr(L, "ref_process_polygon"):  # (a library built before the renderers were added lacks them)
            L.ref_process_polygon.argtypes = [_fp, ctypes.c_float, ctypes.c_int, ctypes.c_int, _fp, _fp, _u32p, _u32p, _u32p]
            L.ref_bitmap.argtypes = [_fp, _fp, ctypes.c_float, ctypes.c_int, ctypes.c_int, _u8p]
            L.ref_matplotlib_slice.argtypes = [_fp, _fp, ctypes.c_float, ctypes.c_int, ctypes.c_int, _fp]
            L.ref_ray_caster.argtypes = [_fp] * 5 + [ctypes.c_float] * 5 + [ctypes.c_uint, ctypes.c_int, ctypes.c_int, _u8p]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


def grid_eval(words, corner, step, dims, x_offset=0):
    """x_offset is emulated by shifting the corner (fp32), which is exact only when
    step * x_offset is representable; validation runs use x_offset = 0."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    if x_offset:
        c[0] = np.float32(c[0] + np.float32(step) * np.float32(x_offset))
    nx, ny, nz = (int(d) for d in dims)
    out = np.empty((nx, ny, nz, 4), np.float32)
    lib().ref_grid_eval(_p(w), _p(c), np.float32(step), nx, ny, nz, _p(out))
    return out


def grid_eval_pymcubes(words, corner, step, dims):
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    out = np.empty((ny, nx, nz), np.float32)
    lib().ref_grid_eval_pymcubes(_p(w), _p(c), np.float32(step), nx, ny, nz, _p(out))
    return out


def subdivision_step(words, corner, step, threshold, dims):
    """-> uint8 [count][4], sorted into INDEX3 order (the reference's atomic order is arbitrary)."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    counter = np.zeros(1, np.uint32)
    lst = np.zeros((nx * ny * nz, 4), np.uint8)
    lib().ref_subdivision_step(_p(w), _p(c), np.float32(step), np.float32(threshold), nx, ny, nz,
                               _p(counter, _u32p), _p(lst, _u8p))
    lst = lst[: int(counter[0])]
    order = np.lexsort((lst[:, 2], lst[:, 1], lst[:, 0]))
    return lst[order].copy()


def mass_properties_step(words, corner, step, threshold, dims):
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    counter = np.zeros(1, np.uint32)
    sums = np.zeros(10, np.uint32)
    lst = np.zeros((nx * ny * nz, 4), np.uint8)
    lib().ref_mass_properties(_p(w), _p(c), np.float32(step), np.float32(threshold), nx, ny, nz,
                              _p(sums, _u32p), _p(counter, _u32p), _p(lst, _u8p))
    lst = lst[: int(counter[0])]
    order = np.lexsort((lst[:, 2], lst[:, 1], lst[:, 0]))
    return sums, lst[order].copy()


def has_renderers():
    return available() and hasattr(lib(), "ref_process_polygon")


def process_polygon(box_corner, step, corners):
    """rendering/polygon2d.cl process_polygon, same call shape as oracle.process_polygon; `starts` in the
    order of a sequential x, y, triangle sweep."""
    corners = _f32(corners)
    gx, gy = corners.shape[0], corners.shape[1]
    cx, cy = gx - 1, gy - 1
    c = _f32(box_corner)[:2].copy()
    vertices = np.zeros((2 * cx * cy, 2), np.float32)
    links = np.zeros(2 * cx * cy, np.uint32)
    starts = np.zeros(2 * cx * cy + 1, np.uint32)
    counter = np.zeros(1, np.uint32)
    lib().ref_process_polygon(_p(c), np.float32(step), cx, cy, _p(corners), _p(vertices), _p(links, _u32p),
                              _p(starts, _u32p), _p(counter, _u32p))
    return vertices, links, starts[: int(counter[0])].copy()


def bitmap(words, origin, step, size):
    """rendering/bitmap.cl over (w, h) work-items -> uint8 [w][h][3] (INDEX2 order, like the kernel)."""
    w = _f32(words)
    out = np.zeros((size[0], size[1], 3), np.uint8)
    lib().ref_bitmap(_p(w), _p(_f32(origin)), np.float32(step), size[0], size[1], _p(out, _u8p))
    return out


def matplotlib_slice(words, corner, step, size):
    """rendering/matplotlib_slice.cl -> float32 [h][w][3] = (distance, gradient x, gradient y)."""
    w = _f32(words)
    out = np.zeros((size[1], size[0], 3), np.float32)
    lib().ref_matplotlib_slice(_p(w), _p(_f32(corner)), np.float32(step), size[0], size[1], _p(out))
    return out


def ray_caster(words, origin, forward, up, right, pixel_tolerance, box_radius, min_distance, max_distance, floor_z,
               options, size):
    """rendering/ray_caster.cl with the kernel's own arguments -> uint8 [w][h][3] (INDEX2 order)."""
    w = _f32(words)
    out = np.zeros((size[0], size[1], 3), np.uint8)
    f = lambda v: _p(_f32(v))  # noqa: E731
    lib().ref_ray_caster(_p(w), f(origin), f(forward), f(up), f(right), np.float32(pixel_tolerance), np.float32(box_radius),
                         np.float32(min_distance), np.float32(max_distance), np.float32(floor_z), int(options),
                         size[0], size[1], _p(out, _u8p))
    return out
