"""Dense-grid evaluation: the reference's `grid_eval` / `grid_eval_pymcubes` kernels
(/root/reference/codecad/grid_eval.cl:2-34; the reference's grid_eval.py only registers
the .cl).  Besides the `.k.` kernel proxy (cl_util) this module offers the call a user
of a dense grid actually wants: evaluate a (slab of a) grid into host memory.
"""
import ctypes

import numpy as np

from . import _lib
from .cl_util.buffer import Buffer
from .geometry import FLOAT4
from .nodes import make_program_buffer


def slab_range(nx, rank, world):
    """x-range [x0, x1) of the slab owned by `rank` when nx planes are split over `world`
    ranks (contiguous in the INDEX3 layout, cl_util/indexing.h:4)."""
    base, rem = divmod(int(nx), int(world))
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


def balanced_slabs(shape, corner, step, dims, world):
    """x-ranges [(x0, x1), ...] for `world` ranks with about equal WORK instead of equal plane counts: with part culling
    the rim of an assembly's grid is cheaper than its middle (cc_grid_eval_cost_profile: the brick masks' estimate per
    layer of eight planes).  Cuts lie on multiples of 8 planes; every rank computes the same list; the results of
    grid_eval do not depend on it.  A program without parts gets slab_range()'s equal slabs."""
    import ctypes
    program = make_program_buffer(shape)
    nx, ny, nz = (int(d) for d in dims)
    world = int(world)
    n_layers = (nx + 7) // 8
    if world <= 1 or n_layers < 2 * world:
        return [slab_range(nx, r, world) for r in range(world)]
    cost = (ctypes.c_double * n_layers)()
    _lib.check(_lib.lib().cc_grid_eval_cost_profile(program.handle, _lib.f3(corner), float(np.float32(step)), nx, ny, nz, 0, cost, n_layers))
    cuts = cut_equal_work(np.array(cost[:], dtype=np.float64), world)
    if cuts is None:
        return [slab_range(nx, r, world) for r in range(world)]
    return [(8 * cuts[r], min(8 * cuts[r + 1], nx)) for r in range(world)]


def cut_equal_work(cost, world):
    """Indices 0 = c_0 < c_1 < ... < c_world = len(cost) that split `cost` into `world` runs of about equal sum, every
    run non-empty; None if the costs are all the same (nothing to balance)."""
    c = np.maximum(np.asarray(cost, dtype=np.float64), 1e-9)
    n = len(c)
    if n < world or np.all(c == c[0]):
        return None
    prefix = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for r in range(1, world):
        target = prefix[-1] * r / world
        k = int(np.searchsorted(prefix, target))
        if k > 0 and abs(prefix[k - 1] - target) < abs(prefix[k] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1] + 1), n - (world - r)))
    cuts.append(n)
    return cuts


def _eval(shape, corner, step, dims, layout, x_offset, out, to_host):
    program = make_program_buffer(shape)
    nx, ny, nz = (int(d) for d in dims)
    if to_host:
        if layout == _lib.LAYOUT_INDEX3_FLOAT4:
            want_shape, dtype = (nx, ny, nz), FLOAT4
        else:
            want_shape, dtype = (ny, nx, nz), np.float32
        if out is None:
            # pinned host array without a device twin; the array's buffer owns the allocation and
            # returns it to the library's page-locked pool when the last view is collected
            from .cl_util.buffer import _Pinned
            out = _Pinned(nx * ny * nz * np.dtype(dtype).itemsize).array(dtype, want_shape)
        if out.nbytes < nx * ny * nz * np.dtype(dtype).itemsize:
            raise RuntimeError("Not enough space to store the grid")
        _lib.check(_lib.lib().cc_grid_eval_to_host(program.handle, _lib.f3(corner), float(np.float32(step)),
                                                   nx, ny, nz, int(x_offset), layout, out.ctypes.data))
        return out
    if out.size < nx * ny * nz * (16 if layout == _lib.LAYOUT_INDEX3_FLOAT4 else 4):
        raise RuntimeError("Output buffer too small for the grid")
    _lib.check(_lib.lib().cc_grid_eval(program.handle, _lib.f3(corner), float(np.float32(step)),
                                       nx, ny, nz, int(x_offset), layout, out.device_ptr, None))
    return out


def release_host_grid(arr):
    """Deprecated no-op: an array returned by grid_eval(out=None) owns its page-locked memory and
    gives it back when it is garbage collected."""


def grid_eval(shape, corner, step, dims, x_offset=0, out=None, device_out=None):
    """Evaluate `shape` on the grid corner + step*(x + x_offset, y, z), x < dims[0] ...

    Returns a host array [nx][ny][nz] of float4 (grad.x, grad.y, grad.z, distance) — the
    reference kernel's INDEX3 layout.  `out` may be a pre-allocated (ideally pinned) host
    array.  With `device_out` (a cl_util.Buffer) the result stays on the device."""
    if device_out is not None:
        return _eval(shape, corner, step, dims, _lib.LAYOUT_INDEX3_FLOAT4, x_offset, device_out, False)
    return _eval(shape, corner, step, dims, _lib.LAYOUT_INDEX3_FLOAT4, x_offset, out, True)


def grid_eval_pymcubes(shape, corner, step, dims, out=None, device_out=None):
    """Distance-only grid in the y-flipped layout PyMCubes wants (grid_eval.cl:16-18):
    host array [ny (flipped)][nx][nz] of float32."""
    if device_out is not None:
        return _eval(shape, corner, step, dims, _lib.LAYOUT_PYMCUBES_FLOAT, 0, device_out, False)
    return _eval(shape, corner, step, dims, _lib.LAYOUT_PYMCUBES_FLOAT, 0, out, True)


def evaluate_points(shape, points):
    """evaluate() at arbitrary points: points [n][3] (or [n][4]) -> float32 [n][4] = (gradient, distance).
    Host arrays in and out."""
    program = make_program_buffer(shape)
    pts = np.asarray(points, dtype=np.float32)
    n = len(pts)
    p4 = np.zeros((n, 4), np.float32)
    p4[:, :3] = pts[:, :3]
    out = np.empty((n, 4), np.float32)
    if n == 0:
        return out
    L = _lib.lib()
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(n * 16, ctypes.byref(d_in)))
    _lib.check(L.cc_buffer_alloc(n * 16, ctypes.byref(d_out)))
    try:
        _lib.check(L.cc_memcpy_h2d_async(d_in, p4.ctypes.data, n * 16, None))
        _lib.check(L.cc_evaluate_points(program.handle, d_in, n, d_out, None))
        _lib.check(L.cc_memcpy_d2h_async(out.ctypes.data, d_out, n * 16, None))
        _lib.check(L.cc_synchronize())
    finally:
        L.cc_buffer_free(d_in)
        L.cc_buffer_free(d_out)
    return out


def synchronize():
    _lib.check(_lib.lib().cc_synchronize())
