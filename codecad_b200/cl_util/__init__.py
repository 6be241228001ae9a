"""Replacement for the reference's `codecad/cl_util` package: same names, CUDA underneath.

    reference symbol                                    here
    cl_util.opencl_manager  (opencl_manager.py:87-144)  manager.OpenCLManager instance
    cl_util.Buffer / BufferList (cl_buffer.py:9-158)    buffer.Buffer / BufferList
    cl_util.interleave (cl_buffer.py:173-198)           pipeline.interleave
    cl_util.interleave2 (cl_util/__init__.py:8-67)      pipeline.interleave2
    cl_util.AssertBuffer / OpenClAssertionError         buffer.AssertBuffer (inert: the CUDA
        (cl_assert.py:8-92)                             kernels have no device-side asserts)

`opencl_manager.k.<kernel>(global_size, local_size, *args, wait_for=None) -> Event` accepts
the four hot-path kernels with the argument lists of the .cl files.
"""
from .buffer import AssertBuffer, Buffer, BufferList, OpenClAssertionError, ProgramBuffer  # noqa: F401
from .manager import Event, instance as opencl_manager  # noqa: F401
from .pipeline import interleave, interleave2  # noqa: F401
from . import parallel_sum  # noqa: F401
