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
