"""Placeholder for cl_util/parallel_sum.py.  The reference generates OpenCL work-group
sum / prefix-sum helpers there (parallel_sum.py:4-74, parallel_sum.cl:1-24) that no
production kernel uses; the CUDA kernels do their scans with warp
shuffles and a decoupled look-back instead.  Kept so `from . import parallel_sum` and
callers of `generate_sum_helper` keep importing."""


def generate_sum_helper(h_file, c_file, type_name, op="a + b", name=None):
    """Accepted and ignored: there is no OpenCL program to append to."""
    return None
