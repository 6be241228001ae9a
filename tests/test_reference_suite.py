"""The reference's OWN pytest suite, unmodified, against the CUDA path.

baseline/_ref/ holds the reference's Python half and its tests (tools/install_reference.py copies
them there in the build container; the directory is git-ignored and travels to the GPU box like
the built libraries).  A child process installs the drop-in (codecad_b200.dropin.load(): our
cl_util / grid_eval / subdivision / mass_properties / rendering modules under the reference's
names), registers host-side counterparts of the kernels the reference's tests bring along
(tests/reference_kernels.py) and runs the reference's test files with pytest.

Deselected, each for a stated reason (nothing else is skipped):
  test_clutil.py::test_assert_*, test_format_c_string_literal, test_sum, test_indexing_prefix_sum
      exercise OpenCL-C device helpers (assert.cl, parallel_sum.cl, run-time compiled source) that
      no production kernel uses; there is no OpenCL compiler behind the drop-in.
"""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

RUNNER = r"""
import os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, os.path.join(root, "tests"), ref, os.path.join(ref, "reference_tests"), os.path.join(ref, "examples"),
                os.path.join(root, "tools", "refstub")]       # refstub: `flags` and a `trimesh` stand-in
import codecad_b200.dropin as dropin
codecad = dropin.load()
import numpy as np
import codecad_b200
import reference_kernels

def evaluate_words(program_buffer, points):
    import ctypes
    from codecad_b200 import _lib
    L = _lib.lib()
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n = len(pts)
    p4 = np.zeros((n, 4), np.float32); p4[:, :3] = pts
    out = np.empty((n, 4), np.float32)
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(n * 16, ctypes.byref(d_in))); _lib.check(L.cc_buffer_alloc(n * 16, ctypes.byref(d_out)))
    _lib.check(L.cc_memcpy_h2d_async(d_in, p4.ctypes.data, n * 16, None))
    _lib.check(L.cc_evaluate_points(program_buffer.handle, d_in, n, d_out, None))
    _lib.check(L.cc_memcpy_d2h_async(out.ctypes.data, d_out, n * 16, None))
    _lib.check(L.cc_synchronize())
    L.cc_buffer_free(d_in); L.cc_buffer_free(d_out)
    return out

reference_kernels.register(codecad.cl_util.opencl_manager, evaluate_words)
import pytest
os.chdir(os.path.join(ref, "reference_tests"))
sys.exit(pytest.main(sys.argv[3:]))
"""

FILES = ["test_mass_properties.py", "test_subdivision.py", "test_dsdf.py", "test_image.py", "test_mesh.py",
         "test_assembly.py", "test_clutil.py", "test_simple2d.py", "test_simple3d.py", "test_shapes.py",
         "test_polygons2d.py", "test_geometry.py", "test_util.py", "test_test_tools.py"]
DESELECT = ["test_clutil.py::test_assert_pass", "test_clutil.py::test_assert_fail", "test_clutil.py::test_assert_fail_multiple",
            "test_clutil.py::test_assert_chaining_pass", "test_clutil.py::test_assert_chaining_multiple_fail",
            "test_clutil.py::test_format_c_string_literal", "test_clutil.py::test_sum", "test_clutil.py::test_indexing_prefix_sum"]


def run_reference_tests(extra=()):
    args = [sys.executable, "-c", RUNNER, ROOT, REF, "-q", "-p", "no:cacheprovider", "-x" if False else "-rfE"]
    for d in DESELECT:
        args += ["--deselect", d]
    args += list(extra) + FILES
    return subprocess.run(args, capture_output=True, text=True, timeout=3000)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "reference_tests")),
                    reason="baseline/_ref not installed (tools/install_reference.py, build container only)")
def test_reference_suite_passes_on_the_cuda_path():
    r = run_reference_tests()
    tail = (r.stdout + r.stderr)[-6000:]
    m = re.search(r"(\d+) passed", r.stdout)
    passed = int(m.group(1)) if m else 0
    failed = re.search(r"(\d+) failed", r.stdout)
    errors = re.search(r"(\d+) error", r.stdout)
    print(tail)
    assert r.returncode == 0 and not failed and not errors, tail
    # the device-dependent part alone is > 250 cases: 9 mass-property, 2 subdivision, ~130 dsdf, 32 images, 6 meshes
    assert passed >= 480, tail   # 487 on the B200 (64 skipped and 1 xfailed by the reference itself)
