"""Drop-in installation into an unmodified reference checkout.

    import codecad_b200.dropin as dropin
    codecad = dropin.load()     # = install() + `import codecad` (the reference checkout on sys.path)
    codecad.mass_properties(shape, 0.1)        # -> runs on the B200 through libcodecad_b200

What it does (INTEGRATION.md has the file-level view):

  * registers `pyopencl` / `pyopencl.cltypes` stand-ins — only the names the reference's
    *Python* touches (mem_flags, cltypes dtypes, Buffer for nodes/program.py:79-84); there is
    no OpenCL behind them, the data path is CUDA;
  * pre-seeds sys.modules so that the reference's own
        codecad.cl_util          (cl_util/__init__.py, opencl_manager.py, cl_buffer.py ...)
        codecad.grid_eval        (grid_eval.py:1-3)
        codecad.subdivision      (subdivision.py)
        codecad.mass_properties  (mass_properties.py)
        codecad.rendering.mesh, codecad.rendering.stl_renderer   (mesh export)
        codecad.rendering.ray_caster, .bitmap, .image, .polygon2d (pictures, 2-D outlines)
    are never loaded: the modules of this package take their place under those names.

Everything else of the reference (shapes, nodes, util, assemblies, rendering front ends)
is imported unchanged.  All eight production kernels of the reference have CUDA counterparts behind
`opencl_manager.k`; any other kernel name raises AttributeError (no OpenCL fallback).
"""
import importlib
import sys
import types

import numpy as np


def _make_pyopencl():
    from .cl_util.buffer import ProgramBuffer
    from .cl_util.manager import Event
    from .geometry import FLOAT2, FLOAT4, UCHAR4

    cl = types.ModuleType("pyopencl")
    cl.__doc__ = "codecad_b200 stand-in for pyopencl (names only; the data path is CUDA)"

    class _Flags:
        def __init__(self, names):
            for i, n in enumerate(names):
                setattr(self, n, 1 << i)

    cl.mem_flags = _Flags(["READ_WRITE", "WRITE_ONLY", "READ_ONLY", "USE_HOST_PTR", "ALLOC_HOST_PTR",
                           "COPY_HOST_PTR", "HOST_WRITE_ONLY", "HOST_READ_ONLY", "HOST_NO_ACCESS"])
    cl.map_flags = _Flags(["READ", "WRITE", "WRITE_INVALIDATE_REGION"])
    cl.command_queue_properties = _Flags(["OUT_OF_ORDER_EXEC_MODE_ENABLE", "PROFILING_ENABLE"])

    def Buffer(context, flags, size=0, hostbuf=None):
        """nodes/program.py:79-84 is the one place the reference constructs a raw
        pyopencl.Buffer on the hot path: a read-only copy of the program words."""
        if hostbuf is None:
            raise RuntimeError("pyopencl stand-in: only program buffers (hostbuf=words) are supported")
        return ProgramBuffer(np.asarray(hostbuf, dtype=np.float32))

    cl.Buffer = Buffer
    cl.Event = Event

    def _no_opencl(*a, **k):
        raise RuntimeError("pyopencl stand-in: there is no OpenCL runtime; the hot path runs in libcodecad_b200")

    cl.create_some_context = _no_opencl
    cl.CommandQueue = _no_opencl
    cl.Program = _no_opencl
    cl.enqueue_copy = _no_opencl
    cl.enqueue_map_buffer = _no_opencl

    ct = types.ModuleType("pyopencl.cltypes")
    # scalar and vector types with pyopencl's layouts: field names x,y,z,w then s4.., a 3-vector
    # occupies the space of a 4-vector
    scalars = {"char": np.int8, "uchar": np.uint8, "short": np.int16, "ushort": np.uint16, "int": np.int32,
               "uint": np.uint32, "long": np.int64, "ulong": np.uint64, "half": np.float16, "float": np.float32,
               "double": np.float64}
    field_names = ["x", "y", "z", "w"] + ["s%d" % i for i in range(4, 16)]
    for sname, stype in scalars.items():
        setattr(ct, sname, stype)
        for count in (2, 3, 4, 8, 16):
            padded = 4 if count == 3 else count
            names = field_names[:count] if count <= 4 else ["s%d" % i for i in range(count)]
            names = names + ["padding%d" % i for i in range(padded - count)]
            dt = np.dtype([(n, stype) for n in names])
            setattr(ct, "%s%d" % (sname, count), dt)

            def filled(value, dt=dt):
                return np.array(tuple([value] * len(dt.names)), dtype=dt)
            setattr(ct, "filled_%s%d" % (sname, count), filled)
    ct.float2 = FLOAT2
    ct.float3 = FLOAT4  # an OpenCL float3 occupies 16 bytes
    ct.float4 = FLOAT4
    ct.uchar4 = UCHAR4
    cl.cltypes = ct
    return cl, ct


def install(force_pyopencl=True):
    """Make `import codecad` (the reference) use this package for its OpenCL layer."""
    if "codecad" in sys.modules and getattr(sys.modules["codecad"], "__codecad_b200__", False) is False \
            and "codecad.cl_util" in sys.modules and sys.modules["codecad.cl_util"].__name__ != "codecad_b200.cl_util":
        raise RuntimeError("codecad is already imported with its own cl_util; call install() first")
    if force_pyopencl or "pyopencl" not in sys.modules:
        cl, ct = _make_pyopencl()
        sys.modules["pyopencl"] = cl
        sys.modules["pyopencl.cltypes"] = ct
    # (the package re-exports functions named like these submodules, so fetch the modules)
    cl_util = importlib.import_module("codecad_b200.cl_util")
    grid_eval = importlib.import_module("codecad_b200.grid_eval")
    mass_properties = importlib.import_module("codecad_b200.mass_properties")
    subdivision = importlib.import_module("codecad_b200.subdivision")
    aliases = {
        "codecad.cl_util": cl_util,
        "codecad.cl_util.opencl_manager": importlib.import_module("codecad_b200.cl_util.manager"),
        "codecad.cl_util.cl_buffer": importlib.import_module("codecad_b200.cl_util.buffer"),
        "codecad.cl_util.cl_assert": importlib.import_module("codecad_b200.cl_util.buffer"),
        "codecad.cl_util.parallel_sum": importlib.import_module("codecad_b200.cl_util.parallel_sum"),
        "codecad.grid_eval": grid_eval,
        "codecad.subdivision": subdivision,
        "codecad.mass_properties": mass_properties,
        # mesh export (rendering/mesh.py imports mcubes + pyopencl, rendering/stl_renderer.py numpy-stl)
        "codecad.rendering.mesh": importlib.import_module("codecad_b200.rendering.mesh"),
        "codecad.rendering.stl_renderer": importlib.import_module("codecad_b200.rendering.stl_renderer"),
        # image renderers and 2-D outlines (their kernels live in libcodecad_b200 too)
        "codecad.rendering.ray_caster": importlib.import_module("codecad_b200.rendering.ray_caster"),
        "codecad.rendering.bitmap": importlib.import_module("codecad_b200.rendering.bitmap"),
        "codecad.rendering.image": importlib.import_module("codecad_b200.rendering.image"),
        "codecad.rendering.polygon2d": importlib.import_module("codecad_b200.rendering.polygon2d"),
        "codecad.rendering.matplotlib_slice": importlib.import_module("codecad_b200.rendering.matplotlib_slice"),
    }
    sys.modules.update(aliases)
    return sorted(aliases)


def load():
    """install(), import the reference's `codecad` package and bind the replaced submodules
    as attributes of it (the import system only does that for modules it loaded itself)."""
    names = install()
    codecad = importlib.import_module("codecad")
    rendering = importlib.import_module("codecad.rendering")
    for full in names:
        parent, _, child = full.rpartition(".")
        mod = sys.modules[full]
        if parent == "codecad.rendering":
            setattr(rendering, child, mod)  # `import codecad.rendering.mesh; codecad.rendering.mesh.f()`
            continue
        if parent != "codecad":
            continue  # cl_util.opencl_manager must stay the *instance* (cl_util/__init__.py:4)
        if child == "mass_properties":
            continue  # codecad.mass_properties is the function (codecad/__init__.py:11)
        setattr(codecad, child, mod)
    codecad.__codecad_b200__ = True
    return codecad
