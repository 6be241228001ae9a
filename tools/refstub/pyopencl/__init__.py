"""Import-time stand-in for pyopencl, used ONLY by tools/ scripts that import the
reference's *Python* half (shape API -> float32 program words) in a container that
has no OpenCL runtime.  It is never imported by the product package and performs no
computation: every device entry point raises.

Surface needed by the reference at import time:
codecad/cl_util/opencl_manager.py:89-98, cl_buffer.py:9-40, nodes/program.py:79-84.
"""
import numpy as _np


class _Bag:
    """Attribute bag: any attribute reads as an int flag."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return 1


mem_flags = _Bag()
map_flags = _Bag()
command_queue_properties = _Bag()


class _Context:
    devices = []


def create_some_context(*a, **k):
    return _Context()


class CommandQueue:
    def __init__(self, context, properties=None):
        self.context = context


class Buffer:
    def __init__(self, context, flags, size=0, hostbuf=None):
        self.size = size if hostbuf is None else hostbuf.nbytes
        self.hostbuf = None if hostbuf is None else _np.array(hostbuf, copy=True)


class Program:
    def __init__(self, *a, **k):
        raise RuntimeError("refstub pyopencl: no OpenCL runtime in this container")


class Event:
    def wait(self):
        pass


def enqueue_copy(*a, **k):
    raise RuntimeError("refstub pyopencl: no OpenCL runtime in this container")


def enqueue_map_buffer(*a, **k):
    raise RuntimeError("refstub pyopencl: no OpenCL runtime in this container")


from . import cltypes  # noqa: E402,F401
