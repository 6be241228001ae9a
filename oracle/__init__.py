"""TEST INFRASTRUCTURE — ctypes front end of the CPU oracle (oracle/sdf_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product package codecad_b200 never does.

`build()` compiles liboracle.so in place (gcc, OpenMP, explicit-FMA arithmetic, no
implicit contraction); the built file travels to the GPU box with the repo snapshot.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_SRC = [os.path.join(_HERE, "sdf_oracle.c"), os.path.join(_HERE, "cc_math_ref.h")]

# -mavx2 -mfma: hardware FMA so fmaf() is one instruction (x86-64-v3; any host that
# carries a B200 has it).  -ffp-contract=off: nothing is fused behind our back.
CFLAGS = ["-O3", "-mavx2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
          "-shared", "-fPIC", "-Wall", "-Wextra"]


def build(force=False):
    if not force and os.path.exists(_SO) and all(
        os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRC
    ):
        return _SO
    cmd = ["gcc"] + CFLAGS + ["-o", _SO, _SRC[0], "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None
_fp = ctypes.POINTER(ctypes.c_float)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_validate.argtypes = [_fp, ctypes.c_int]
        L.oracle_evaluate_points.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_long, _fp]
        L.oracle_grid_eval.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_float] + [ctypes.c_int] * 4 + [_fp]
        L.oracle_grid_eval_pymcubes.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_float] + [ctypes.c_int] * 3 + [_fp]
        L.oracle_subdivision_step.argtypes = (
            [_fp, ctypes.c_int, _fp, ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 3 + [_u32p, _u8p]
        )
        L.oracle_mass_properties_step.argtypes = (
            [_fp, ctypes.c_int, _fp, ctypes.c_float, ctypes.c_float] + [ctypes.c_int] * 3 + [_u32p, _u32p, _u8p]
        )
        L.oracle_process_polygon.argtypes = [_fp, ctypes.c_float, ctypes.c_int, ctypes.c_int, _fp, _fp, _u32p, _u32p, _u32p]
        L.oracle_executed_flops.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_float] + [ctypes.c_int] * 5 + [
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
        L.oracle_math_probe.argtypes = [ctypes.c_int, _fp, _fp, ctypes.c_long, _fp, _fp]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


def _check(rc):
    if rc < 0:
        raise ValueError("oracle: malformed program (code %d)" % rc)
    return rc


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def validate(words):
    w = _f32(words)
    return _check(lib().oracle_validate(_p(w), len(w)))


def evaluate_points(words, pts):
    w = _f32(words)
    pts = _f32(pts).reshape(-1, 3)
    out = np.empty((len(pts), 4), np.float32)
    _check(lib().oracle_evaluate_points(_p(w), len(w), _p(pts), len(pts), _p(out)))
    return out


def grid_eval(words, corner, step, dims, x_offset=0):
    """float4 grid in the reference's INDEX3 layout: out[x][y][z] = (grad.xyz, dist)."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    out = np.empty((nx, ny, nz, 4), np.float32)
    _check(lib().oracle_grid_eval(_p(w), len(w), _p(c), np.float32(step), nx, ny, nz, int(x_offset), _p(out)))
    return out


def grid_eval_pymcubes(words, corner, step, dims):
    """distance-only grid in the PyMCubes layout of grid_eval.cl:18: flat index
    z + (x + (ny-1-y)*nx)*nz, i.e. array [ny (flipped)][nx][nz]."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    out = np.empty((ny, nx, nz), np.float32)
    _check(lib().oracle_grid_eval_pymcubes(_p(w), len(w), _p(c), np.float32(step), nx, ny, nz, _p(out)))
    return out


def subdivision_step(words, corner, step, threshold, dims):
    """-> uint8 array [count][4] of (x, y, z, 0) in INDEX3 order."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    counter = np.zeros(1, np.uint32)
    lst = np.zeros((nx * ny * nz, 4), np.uint8)
    _check(lib().oracle_subdivision_step(_p(w), len(w), _p(c), np.float32(step), np.float32(threshold),
                                         nx, ny, nz, _p(counter, _u32p), _p(lst, _u8p)))
    return lst[: int(counter[0])].copy()


def mass_properties_step(words, corner, step, threshold, dims):
    """-> (sums uint32[10] in order xx,xy,xz,x,yy,yz,y,zz,z,n ; list uint8[count][4])."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    counter = np.zeros(1, np.uint32)
    sums = np.zeros(10, np.uint32)
    lst = np.zeros((nx * ny * nz, 4), np.uint8)
    _check(lib().oracle_mass_properties_step(_p(w), len(w), _p(c), np.float32(step), np.float32(threshold),
                                             nx, ny, nz, _p(sums, _u32p), _p(counter, _u32p), _p(lst, _u8p)))
    return sums, lst[: int(counter[0])].copy()


def process_polygon(box_corner, step, corners):
    """rendering/polygon2d.cl process_polygon over the float4 grid corners[gx][gy][4] (grid_eval output
    with nz = 1 squeezed).  -> (vertices float32 [cells][2], links uint32 [cells], starts uint32 [count])
    with cells = 2*(gx-1)*(gy-1); vertices of triangles the outline does not cross stay zero."""
    corners = _f32(corners)
    gx, gy = corners.shape[0], corners.shape[1]
    cx, cy = gx - 1, gy - 1
    c = _f32(box_corner)[:2].copy()
    vertices = np.zeros((2 * cx * cy, 2), np.float32)
    links = np.zeros(2 * cx * cy, np.uint32)
    starts = np.zeros(2 * cx * cy + 1, np.uint32)
    counter = np.zeros(1, np.uint32)
    _check(lib().oracle_process_polygon(_p(c), np.float32(step), cx, cy, _p(corners), _p(vertices),
                                        _p(links, _u32p), _p(starts, _u32p), _p(counter, _u32p)))
    return vertices, links, starts[: int(counter[0])].copy()


def executed_flops(words, corner, step, dims, stride=16, x_offset=0):
    """Mean executed-branch algorithmic flop/point (SURVEY.md 8(a3) counting rules, the reference's
    formulation of every op) over every `stride`-th point per axis of the grid -> (mean, points)."""
    w = _f32(words)
    c = _f32(corner)[:3].copy()
    nx, ny, nz = (int(d) for d in dims)
    mean, pts = ctypes.c_double(), ctypes.c_longlong()
    _check(lib().oracle_executed_flops(_p(w), len(w), _p(c), np.float32(step), nx, ny, nz, int(x_offset), int(stride),
                                       ctypes.byref(mean), ctypes.byref(pts)))
    return float(mean.value), int(pts.value)


def math_probe(which, a, b=None):
    a = _f32(a).ravel()
    b = _f32(np.zeros_like(a) if b is None else b).ravel()
    out = np.empty_like(a)
    out2 = np.empty_like(a)
    lib().oracle_math_probe(int(which), _p(a), _p(b), len(a), _p(out), _p(out2))
    return out, out2
