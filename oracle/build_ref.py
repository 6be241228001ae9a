"""TEST INFRASTRUCTURE — build oracle/_ref/libcodecad_ref.so: the reference's OWN OpenCL
device sources, compiled for the host CPU.

No OpenCL runtime exists in this image, so the reference cannot run as shipped.  Its
device code, however, is ~600 lines of OpenCL C that is almost C++.  This recipe

  1. imports the reference's Python (with tools/refstub standing in for pyopencl) and
     collects the exact program text its OpenCLManager would compile — common header +
     compile units in registration order (cl_util/opencl_manager.py:116-127), including
     the `evaluate()` interpreter that nodes/codegen.py GENERATES;
  2. keeps the units of the hot path (op library, evaluate, grid_eval, subdivision_step,
     mass_properties) and drops rendering/assert/parallel_sum units;
  3. rewrites the three OpenCL-only syntaxes C++ cannot parse — `(float4)(..)` vector
     literals, unsuffixed float constants (the reference builds with
     -cl-single-precision-constant), address-space qualifiers (macros in opencl_shim.hpp);
  4. compiles the result with g++ inside `namespace clref` against oracle/opencl_shim.hpp,
     plus the NDRange driver loops at the bottom of this file.

The reference sources are read where they lie under /root/reference; nothing of them is
copied into the repository — the translated unit and the .so go to oracle/_ref/ (git-ignored,
but shipped to the GPU box with the snapshot).  On a machine without /root/reference this
is a no-op that keeps an existing prebuilt library.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "_ref")
SO = os.path.join(OUT_DIR, "libcodecad_ref.so")
REF = os.environ.get("CODECAD_REFERENCE", "/root/reference")

KEEP = {"util.h", "indexing.h", "util.cl", "common.h", "common.cl", "simple2d.cl", "simple3d.cl",
        "polygons2d.cl", "unsafe.cl", "gears.cl", "codegen.py", "grid_eval.cl", "subdivision.cl",
        "mass_properties.cl"}

DRIVER = r'''
thread_local ndrange g_nd;
}  // namespace clref

using namespace clref;

// NDRange driver loops: one work-item at a time (work-group size 1), x planes over OpenMP.
extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }

void ref_grid_eval(const float *scene, const float *corner, float step, int nx, int ny, int nz, float *out)
{
    clref::float4 c(corner[0], corner[1], corner[2], 0.0f);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                g_nd = ndrange{{(uint)x, (uint)y, (uint)z}, {(uint)nx, (uint)ny, (uint)nz}};
                grid_eval(scene, c, step, reinterpret_cast<clref::float4 *>(out));
            }
}

void ref_grid_eval_pymcubes(const float *scene, const float *corner, float step, int nx, int ny, int nz, float *out)
{
    clref::float4 c(corner[0], corner[1], corner[2], 0.0f);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                g_nd = ndrange{{(uint)x, (uint)y, (uint)z}, {(uint)nx, (uint)ny, (uint)nz}};
                grid_eval_pymcubes(scene, c, step, out);
            }
}

void ref_subdivision_step(const float *scene, const float *corner, float step, float thr, int nx, int ny, int nz,
                          unsigned *counter, unsigned char *list)
{
    clref::float4 c(corner[0], corner[1], corner[2], 0.0f);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                g_nd = ndrange{{(uint)x, (uint)y, (uint)z}, {(uint)nx, (uint)ny, (uint)nz}};
                subdivision_step(scene, c, step, thr, counter, reinterpret_cast<clref::uchar4 *>(list));
            }
}

void ref_mass_properties(const float *scene, const float *corner, float step, float thr, int nx, int ny, int nz,
                         unsigned *sums, unsigned *counter, unsigned char *list)
{
    clref::float4 c(corner[0], corner[1], corner[2], 0.0f);
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                g_nd = ndrange{{(uint)x, (uint)y, (uint)z}, {(uint)nx, (uint)ny, (uint)nz}};
                mass_properties(scene, c, step, thr, sums, counter, reinterpret_cast<clref::uchar4 *>(list));
            }
}

}  // extern "C"
'''


def collect_reference_program():
    """The OpenCL program text of the reference, unit by unit, as (origin_basename, text)."""
    stub = os.path.join(REPO, "tools", "refstub")
    saved = list(sys.path)
    sys.path[:0] = [stub, REF]
    try:
        import warnings
        warnings.simplefilter("ignore")
        import codecad  # noqa: F401  (registers every compile unit at import time)
        from codecad.cl_util import opencl_manager as mgr
    finally:
        sys.path[:] = saved
    pieces = list(mgr.common_header.pieces)
    for cu in mgr._compile_units:
        pieces.extend(cu.pieces)
    units = []
    origin = None
    for p in pieces:
        m = re.match(r'#line\s+\d+\s+"(.*)"\s*$', p)
        if m:
            origin = os.path.basename(m.group(1).encode().decode("unicode_escape"))
            continue
        units.append((origin, p))
    return units


_VEC = re.compile(r"\(\s*(float2|float3|float4|uint3|uchar4)\s*\)\s*\(")
_FLT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])")


def translate(units):
    out = ['#include "../opencl_shim.hpp"', "#include <omp.h>", "namespace clref {"]
    for origin, text in units:
        if origin not in KEEP:
            continue
        text = _VEC.sub(lambda m: "mk_%s(" % m.group(1), text)
        text = _FLT.sub(lambda m: m.group(1) + "f", text)
        out.append("// ---- from the reference: %s ----" % origin)
        out.append(text)
    out.append(DRIVER)
    return "\n".join(out)


def build(force=False):
    if not os.path.isdir(os.path.join(REF, "codecad")):
        if os.path.exists(SO):
            return SO
        raise RuntimeError("reference checkout not found at %s and no prebuilt %s" % (REF, SO))
    os.makedirs(OUT_DIR, exist_ok=True)
    shim = os.path.join(HERE, "opencl_shim.hpp")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(
            os.path.getmtime(shim), os.path.getmtime(os.path.abspath(__file__))):
        return SO
    src = os.path.join(OUT_DIR, "reference_program.cpp")
    with open(src, "w") as f:
        f.write(translate(collect_reference_program()))
    cmd = ["g++", "-std=c++17", "-O2", "-mavx2", "-mfma", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
           "-fpermissive", "-Wno-narrowing", "-w", "-shared", "-fPIC", "-o", SO, src]
    try:
        subprocess.run(cmd, check=True)
    finally:
        if not os.environ.get("CODECAD_KEEP_REF_TRANSLATION"):
            os.remove(src)  # the translated unit is reference text: never left lying around
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
