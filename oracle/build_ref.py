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
        "mass_properties.cl",
        # the renderers around evaluate(): pin the oracle's restatements of them
        "assert.h", "opencl_manager.py", "ray_caster.cl", "bitmap.cl", "polygon2d.cl", "matplotlib_slice.cl"}
# read straight from the tree: its Python module needs matplotlib to import, so the unit is never registered
EXTRA_FILES = ["codecad/rendering/matplotlib_slice.cl"]

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

// ---- the renderers around evaluate() ----
void ref_process_polygon(const float *corner2, float step, int cx, int cy, const float *corners, float *vertices,
                         unsigned *links, unsigned *starts, unsigned *counter)
{
    clref::float2 c(corner2[0], corner2[1]);
    // sequential: the order of `starts` is then x, y, triangle (the reference's atomic_inc order is arbitrary)
    for (int x = 0; x < cx; ++x)
        for (int y = 0; y < cy; ++y)
            for (int t = 0; t < 2; ++t) {
                g_nd = ndrange{{(uint)x, (uint)y, (uint)t}, {(uint)cx, (uint)cy, 2u}};
                process_polygon(c, step, reinterpret_cast<clref::float4 *>(const_cast<float *>(corners)),
                                reinterpret_cast<clref::float2 *>(vertices), links, starts, counter);
            }
}

void ref_bitmap(const float *scene, const float *origin, float step, int w, int h, unsigned char *out)
{
    clref::float4 o(origin[0], origin[1], origin[2], 0.0f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            g_nd = ndrange{{(uint)x, (uint)y, 0u}, {(uint)w, (uint)h, 1u}};
            bitmap(scene, o, step, out);
        }
}

void ref_matplotlib_slice(const float *scene, const float *corner, float step, int w, int h, float *out)
{
    clref::float4 c(corner[0], corner[1], corner[2], 0.0f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            g_nd = ndrange{{(uint)x, (uint)y, 0u}, {(uint)w, (uint)h, 1u}};
            matplotlib_slice(scene, c, step, out);
        }
}

void ref_ray_caster(const float *scene, const float *origin, const float *forward, const float *up, const float *right,
                    float pixelTolerance, float boxRadius, float minDistance, float maxDistance, float floorZ,
                    unsigned renderOptions, int w, int h, unsigned char *out)
{
    clref::float4 o(origin[0], origin[1], origin[2], 0.0f), f(forward[0], forward[1], forward[2], 0.0f),
        u(up[0], up[1], up[2], 0.0f), r(right[0], right[1], right[2], 0.0f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            g_nd = ndrange{{(uint)x, (uint)y, 0u}, {(uint)w, (uint)h, 1u}};
            ray_caster(scene, o, f, u, r, pixelTolerance, boxRadius, minDistance, maxDistance, floorZ, renderOptions, out,
                       nullptr);
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
    have = {o for o, _ in units}
    for rel in EXTRA_FILES:
        if os.path.basename(rel) not in have:
            with open(os.path.join(REF, rel)) as f:
                units.append((os.path.basename(rel), f.read()))
    return units


_VEC = re.compile(r"\(\s*(float2|float3|float4|uint2|uint3|int2|uchar4)\s*\)\s*\(")
_FLT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])")


def translate(units):
    out = ['#include "../opencl_shim.hpp"', "#include <omp.h>", "namespace clref {",
           "#define ASSERT_BUFFER_SIZE 1024  /* cl_util/cl_assert.py; asserts themselves stay disabled */"]
    for origin, text in units:
        if origin not in KEEP:
            continue
        text = text.replace("const __constant", "__constant")
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
