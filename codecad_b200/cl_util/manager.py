"""`opencl_manager` replacement (cl_util/opencl_manager.py:87-144).

The reference's singleton owns an OpenCL context, an out-of-order queue and a lazily built
program, and exposes kernels as `opencl_manager.k.<name>(global, local, *args, wait_for=)`.
Here the context is the CUDA context inside libcodecad_b200.so (created lazily on first
use, so importing needs no GPU), the "queue" is its in-order compute stream — every
`wait_for` list is therefore trivially satisfied — and the kernel table is fixed.
"""
import ctypes

import numpy as np

from .. import _lib


class Event:
    """pyopencl.Event stand-in: .wait(), .profile.start/.end (ns)."""

    def __init__(self, handle=None):
        self._h = handle

    def wait(self):
        if self._h is not None:
            _lib.check(_lib.lib().cc_event_wait(self._h))
        return self

    def __del__(self):
        try:
            if self._h is not None and _lib._lib is not None:
                _lib._lib.cc_event_destroy(self._h)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def _new_event_ref():
    return ctypes.c_void_p()


class CompileUnit:
    """Inert: shapes/*.py and rendering/*.py register .cl sources at import time
    (opencl_manager.py:21-70); there is nothing to compile here."""

    def __init__(self, *a, **k):
        self.pieces = []
        self.include_origin = True

    def clear(self):
        self.pieces = []

    def code(self, extra_headers=()):
        return ""

    def append_resource(self, resource_name, stacklevel=1):
        self.pieces.append(("resource", resource_name))

    def append_define(self, name, value, stacklevel=1):
        self.pieces.append(("define", name, value))

    def append_flags(self, flags, prefix=None):
        self.pieces.append(("flags", flags))

    def append(self, code, stacklevel=1):
        self.pieces.append(("code", code))


class _Context:
    """opencl_manager.context: only passed back into Buffer constructors."""
    devices = ()


class _Queue:
    def __init__(self, context):
        self.context = context

    def finish(self):
        _lib.check(_lib.lib().cc_synchronize())


def _dims(global_size):
    g = tuple(int(v) for v in global_size)
    return g + (1,) * (3 - len(g))


class _Kernels:
    """`opencl_manager.k`: the hot-path kernels and the renderers around evaluate(), with the .cl argument lists."""

    def __init__(self, manager):
        self.manager = manager

    # grid_eval.cl:23-25   grid_eval(scene, float4 boxCorner, float boxStep, float4* output)
    def grid_eval(self, global_size, local_size, program, corner, step, output, wait_for=None):
        return self._grid(global_size, program, corner, step, output, _lib.LAYOUT_INDEX3_FLOAT4)

    # grid_eval.cl:2-4
    def grid_eval_pymcubes(self, global_size, local_size, program, corner, step, output, wait_for=None):
        return self._grid(global_size, program, corner, step, output, _lib.LAYOUT_PYMCUBES_FLOAT)

    def _grid(self, global_size, program, corner, step, output, layout):
        nx, ny, nz = _dims(global_size)
        elem = 16 if layout == _lib.LAYOUT_INDEX3_FLOAT4 else 4
        if output.size < nx * ny * nz * elem:
            raise RuntimeError("Output buffer too small for the launch")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_grid_eval(program.handle, _lib.f3(corner), float(np.float32(step)),
                                           nx, ny, nz, 0, layout, output.device_ptr, ctypes.byref(ev)))
        return Event(ev)

    # subdivision.cl:12-16
    def subdivision_step(self, global_size, local_size, program, corner, step, threshold, counter, lst,
                         wait_for=None):
        nx, ny, nz = _dims(global_size)
        if lst.size < nx * ny * nz * 4:
            raise RuntimeError("Index list buffer too small for the launch")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_subdivision_step(program.handle, _lib.f3(corner), float(np.float32(step)),
                                                  float(np.float32(threshold)), nx, ny, nz,
                                                  counter.device_ptr, lst.device_ptr, ctypes.byref(ev)))
        return Event(ev)

    # mass_properties.cl:7-12
    def mass_properties(self, global_size, local_size, program, corner, step, threshold, sums, counter, lst,
                        wait_for=None):
        nx, ny, nz = _dims(global_size)
        if lst.size < nx * ny * nz * 4:
            raise RuntimeError("Index list buffer too small for the launch")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_mass_properties_step(program.handle, _lib.f3(corner), float(np.float32(step)),
                                                      float(np.float32(threshold)), nx, ny, nz,
                                                      sums.device_ptr, counter.device_ptr, lst.device_ptr,
                                                      ctypes.byref(ev)))
        return Event(ev)

    # rendering/bitmap.cl:1-3   bitmap(scene, float4 origin, float stepSize, uchar* output)
    def bitmap(self, global_size, local_size, program, origin, step_size, output, wait_for=None):
        w, h, _ = _dims(global_size)
        if output.size < w * h * 3:
            raise RuntimeError("Output buffer too small for the launch")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_bitmap(program.handle, _lib.f3(origin), float(np.float32(step_size)), w, h,
                                        output.device_ptr, ctypes.byref(ev)))
        return Event(ev)

    # rendering/ray_caster.cl:147-156 (the AssertBuffer argument is accepted and ignored)
    def ray_caster(self, global_size, local_size, program, origin, forward, up, right, pixel_tolerance, box_radius,
                   min_distance, max_distance, floor_z, render_options, output, assert_buffer=None, wait_for=None):
        w, h, _ = _dims(global_size)
        if output.size < w * h * 3:
            raise RuntimeError("Output buffer too small for the launch")
        ev = _new_event_ref()
        f = lambda x: float(np.float32(x))  # noqa: E731
        _lib.check(_lib.lib().cc_ray_caster(program.handle, _lib.f3(origin), _lib.f3(forward), _lib.f3(up), _lib.f3(right),
                                            f(pixel_tolerance), f(box_radius), f(min_distance), f(max_distance),
                                            f(floor_z), int(render_options), w, h, output.device_ptr, None,
                                            ctypes.byref(ev)))
        return Event(ev)

    # rendering/matplotlib_slice.cl:1-4   matplotlib_slice(scene, float4 boxCorner, float boxStep, float* output)
    def matplotlib_slice(self, global_size, local_size, program, corner, step, output, wait_for=None):
        w, h, _ = _dims(global_size)
        if output.size < w * h * 12:
            raise RuntimeError("Output buffer too small for the launch")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_matplotlib_slice(program.handle, _lib.f3(corner), float(np.float32(step)), w, h,
                                                  output.device_ptr, ctypes.byref(ev)))
        return Event(ev)

    # rendering/polygon2d.cl:82-90   process_polygon(float2 boxCorner, float boxStep, corners, vertices, links,
    #                                                starts, startCounter)
    def process_polygon(self, global_size, local_size, box_corner, box_step, corners, vertices, links, starts,
                        start_counter, wait_for=None):
        cx, cy, two = _dims(global_size)
        if two != 2:
            raise RuntimeError("process_polygon runs over (cells_x, cells_y, 2)")
        c = np.asarray(box_corner)
        c2 = (ctypes.c_float * 2)(float(c["x"]), float(c["y"])) if c.dtype.names else (ctypes.c_float * 2)(
            *[float(np.float32(v)) for v in np.asarray(c, dtype=np.float64).ravel()[:2]])
        corners_ptr = corners.device_ptr if hasattr(corners, "device_ptr") else corners
        if vertices.size < cx * cy * 2 * 8 or links.size < cx * cy * 2 * 4:
            raise RuntimeError("Output buffer too small for the launch")
        max_starts = min(int(starts.size) // 4, 1024)
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_process_polygon(c2, float(np.float32(box_step)), cx, cy, corners_ptr,
                                                 vertices.device_ptr, links.device_ptr, starts.device_ptr, max_starts,
                                                 start_counter.device_ptr, ctypes.byref(ev)))
        return Event(ev)

    def __getattr__(self, name):
        extra = self.__dict__.get("manager")._extra_kernels if "manager" in self.__dict__ else {}
        if name in extra:
            return extra[name]
        raise AttributeError(
            "kernel %r has no CUDA counterpart (available: grid_eval, grid_eval_pymcubes, subdivision_step, "
            "mass_properties, bitmap, ray_caster, process_polygon, matplotlib_slice); there is no OpenCL fallback" % name)


class OpenCLManager:
    def __init__(self):
        self.context = _Context()
        self.queue = _Queue(self.context)
        self._compile_units = []
        self.common_header = CompileUnit()
        self._extra_kernels = {}
        self.k = _Kernels(self)
        self.max_register_count = 512  # EVAL_REGISTER_COUNT, nodes/__init__.py:6

    def register_kernel(self, name, fn):
        """Make `k.<name>(global_size, local_size, *args, wait_for=None)` call fn.  For callers that
        bring their own kernels built around evaluate() — the reference's tests do (tests/test_dsdf.cl)
        — and express them on the host side of the boundary with codecad_b200.evaluate_points."""
        self._extra_kernels[name] = fn

    def add_compile_unit(self, *args, **kwargs):
        cu = CompileUnit(*args, **kwargs)
        self._compile_units.append(cu)
        return cu

    def get_program(self):
        return self


instance = OpenCLManager()
