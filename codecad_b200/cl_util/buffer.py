"""Device buffers with the interface of the reference's `cl_util.Buffer`
(/root/reference/codecad/cl_util/cl_buffer.py:9-158), backed by libcodecad_b200.

Ownership follows the reference: the Buffer owns a device allocation and an optional host
mirror `.array`; here the mirror is PINNED host memory (the reference asks OpenCL for
ALLOC_HOST_PTR, cl_buffer.py:36-40), so reads and writes are true async DMA.
"""
import contextlib
import ctypes

import numpy as np

from .. import _lib
from ..geometry import FLOAT2, FLOAT4, UCHAR4  # noqa: F401  (re-exported dtypes)
from .manager import Event, _new_event_ref


class _Pinned:
    """numpy view over cudaMallocHost memory; freed when the last view dies."""

    def __init__(self, nbytes):
        self.ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().cc_host_alloc(max(int(nbytes), 1), ctypes.byref(self.ptr)))
        self.nbytes = int(nbytes)

    def array(self, dtype, shape, offset=0):
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr.value)
        buf._pinned_owner = self  # the array keeps the allocation alive through its buffer
        # (like numpy.empty(shape, dtype): a sub-array dtype such as (float4, (2,)) adds its own axes)
        return np.ndarray(shape=shape, dtype=dtype, buffer=buf, offset=offset)

    def __del__(self):
        try:
            if self.ptr and _lib._lib is not None:
                _lib._lib.cc_host_free(self.ptr)
        except Exception:  # noqa: BLE001
            pass


class Buffer:
    @staticmethod
    def dual_dtype(scalar):
        return np.dtype([(name, scalar) for name in "xy"])

    @staticmethod
    def quad_dtype(scalar):
        return np.dtype([(name, scalar) for name in "xyzw"])

    def __init__(self, dtype, shape, mem_flags=None, queue=None):
        self.queue = queue
        self.dtype = np.dtype(dtype)
        self.nitems = 1
        try:
            for s in shape:
                self.nitems *= int(s)
            self.shape = tuple(int(s) for s in shape)
        except TypeError:
            self.nitems = int(shape)
            self.shape = (int(shape),)
        self.size = self.nitems * self.dtype.itemsize
        self.array = None
        self._pinned = None
        self._dptr = ctypes.c_void_p()
        _lib.check(_lib.lib().cc_buffer_alloc(self.size, ctypes.byref(self._dptr)))

    @property
    def device_ptr(self):
        if not self._dptr:
            raise RuntimeError("Buffer has been released")
        return self._dptr

    def release(self):
        if self._dptr and _lib._lib is not None:
            _lib._lib.cc_buffer_free(self._dptr)
        self._dptr = ctypes.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001
            pass

    def create_host_side_array(self):
        self._pinned = _Pinned(self.size)
        self.array = self._pinned.array(self.dtype, self.shape)

    def _process_array(self, array):
        if array is None:
            if self.array is None:
                self.create_host_side_array()
            return self.array
        return array

    def enqueue_read(self, out=None, wait_for=None):
        array = self._process_array(out)
        if array.nbytes < self.size:
            raise RuntimeError("Not enough space to store contents of the buffer")
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_memcpy_d2h_async(array.ctypes.data, self.device_ptr, self.size, ctypes.byref(ev)))
        return Event(ev)

    def read(self, out=None, wait_for=None):
        array = self._process_array(out)
        self.enqueue_read(array).wait()
        return array

    def enqueue_write(self, a=None, wait_for=None):
        array = self._process_array(a)
        array = np.ascontiguousarray(array)
        if array.nbytes > self.size:
            raise RuntimeError("Not enough space to store contents in the buffer")
        ev = _new_event_ref()
        if a is not None:
            # pageable source: the copy must have consumed it before we return
            _lib.check(_lib.lib().cc_memcpy_h2d_async(self.device_ptr, array.ctypes.data, array.nbytes, ctypes.byref(ev)))
            e = Event(ev)
            e.wait()
            return e
        _lib.check(_lib.lib().cc_memcpy_h2d_async(self.device_ptr, array.ctypes.data, array.nbytes, ctypes.byref(ev)))
        return Event(ev)

    def enqueue_zero_fill_compatible(self, wait_for=None):
        ev = _new_event_ref()
        _lib.check(_lib.lib().cc_memset_async(self.device_ptr, 0, self.size, ctypes.byref(ev)))
        return Event(ev)

    @contextlib.contextmanager
    def map(self, map_flags=None, offset=None, shape=None, wait_for=None):
        """Maps (part of) the buffer as a numpy array, like cl_buffer.py:101-122 /
        pyopencl.enqueue_map_buffer: `offset` is in BYTES, `shape` defaults to the whole buffer.  A map
        with READ (or no flags) sees the device contents; the mapped range is written back on exit — also
        when the body raises — only if WRITE or WRITE_INVALIDATE_REGION was asked for (bits 2 and 4 of
        the stand-in map_flags of dropin.py; None = read and write)."""
        flags = 3 if map_flags is None else int(map_flags)
        offset = int(offset or 0)
        if shape is None:
            shape = self.shape
        try:
            shape = tuple(int(v) for v in shape)
        except TypeError:
            shape = (int(shape),)
        if offset % self.dtype.itemsize:
            raise RuntimeError("map offset must be a multiple of the item size")
        first = offset // self.dtype.itemsize
        count = int(np.prod(shape)) if shape else 1
        if first + count > self.nitems:
            raise RuntimeError("mapped range exceeds the buffer")
        writes = bool(flags & 6)
        invalidate = bool(flags & 4) and not (flags & 1)
        self._process_array(None)            # the page-locked mirror (.array)
        if not invalidate:
            self.read()                      # READ / WRITE: the current device contents
        view = self._pinned.array(self.dtype, shape, offset)
        try:
            yield view
        finally:
            if writes:
                nbytes = count * self.dtype.itemsize
                ev = _new_event_ref()
                _lib.check(_lib.lib().cc_memcpy_h2d_async(ctypes.c_void_p(self.device_ptr.value + offset),
                                                          view.ctypes.data, nbytes, ctypes.byref(ev)))
                Event(ev).wait()

    def __getitem__(self, key):
        return self.array[key]

    def __setitem__(self, key, value):
        self.array[key] = value

    def __len__(self):
        return self.nitems


class ProgramBuffer:
    """What `nodes.make_program_buffer(shape)` returns: the reference returns a read-only
    pyopencl.Buffer holding the float32 words (nodes/program.py:79-84); here it is the
    decoded, device-resident program (cc_program)."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, dtype=np.float32)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().cc_program_create(
            self.words.ctypes.data_as(_lib.c_float_p), len(self.words), ctypes.byref(self._h)))
        self.size = self.words.nbytes

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("ProgramBuffer has been released")
        return self._h

    @property
    def info(self):
        info = _lib.ProgramInfo()
        _lib.check(_lib.lib().cc_program_get_info(self.handle, ctypes.byref(info)))
        return info

    SINK_FLOAT4, SINK_PYMCUBES, SINK_CLASSIFY, SINK_MASS = 1, 2, 4, 8
    SINK_RAY, SINK_BITMAP = 16, 32  # image renderers (one ray / pixel per thread)
    SINK_POINTS = 64                # evaluate_points()
    SINK_PARTS = 128                  # dense float4 grids of an assembly: the part-culling pair of kernels
    SINK_COLUMNS = 256                # dense float4 grids of extrusions: the column kernels
    SINK_TILES_PYMCUBES, SINK_TILES_CLASSIFY, SINK_TILES_MASS = 512, 1024, 2048   # the hierarchy sinks of such programs

    def specialize(self, points_per_thread=0, sinks=0):
        """Compile scene-specialised kernels for this program (NVRTC, seconds per sink); later
        launches of those sinks use them.  `sinks`: OR of SINK_* (0 = all four).  Returns the
        compile time in seconds."""
        secs = ctypes.c_double()
        _lib.check(_lib.lib().cc_program_specialize(self.handle, int(points_per_thread), int(sinks),
                                                    ctypes.byref(secs)))
        return secs.value

    def use_specialized(self, enable=True):
        return bool(_lib.lib().cc_program_use_specialized(self.handle, 1 if enable else 0))

    def wait_specialized(self, sinks=0):
        """Block until the background-compiled specialised kernels of `sinks` are loaded (starting
        the compilation if needed).  Returns (number of sinks ready, compile seconds so far)."""
        secs = ctypes.c_double()
        n = _lib.check(_lib.lib().cc_program_specialize_wait(self.handle, int(sinks), ctypes.byref(secs)))
        return n, secs.value

    def microcode(self):
        n = _lib.lib().cc_program_get_microcode(self.handle, None, 0)
        out = np.zeros(n, np.uint32)
        _lib.lib().cc_program_get_microcode(self.handle, out.ctypes.data_as(_lib.c_u32_p), n)
        return out

    def release(self):
        if self._h and _lib._lib is not None:
            _lib._lib.cc_program_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001
            pass


class BufferList:
    """cl_buffer.py:134-158"""

    def __init__(self, buffers=()):
        self.buffers = list(buffers)

    def add(self, buff):
        self.buffers.append(buff)

    def release(self):
        try:
            for buff in self.buffers:
                buff.release()
        finally:
            self.buffers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc_info):
        self.release()


class OpenClAssertionError(Exception):
    """cl_assert.py:8-30 — never raised: the CUDA kernels carry no device-side asserts."""


class AssertBuffer:
    """cl_assert.py:33-92, inert."""

    ASSERT_ENABLED = False

    def __init__(self, queue=None):
        pass

    def reset(self):
        return Event()

    def check(self, wait_for=None):
        return None
