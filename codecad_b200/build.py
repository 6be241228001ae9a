"""Build libcodecad_b200.so in-tree with nvcc for sm_100a.

    python -m codecad_b200.build [--force] [--verbose]

-fmad=false / -ffp-contract=off are part of the numerical contract (csrc/cc_math.cuh):
all fused multiply-adds are explicit.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcodecad_b200.so")
SOURCES = ["cc_kernels.cu", "cc_mesh.cu", "cc_polygon.cu", "cc_program.cpp", "cc_api.cpp", "cc_jit.cpp"]
HEADERS = ["cc_internal.h", "cc_microcode.h", "cc_math.cuh", "cc_ops.cuh", "cc_body.cuh", "cc_render.cuh", "cc_device_types.h", "cc_mc_table.h", "cc_scan.cuh",
           os.path.join("..", "..", "include", "codecad_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-shared",
    "-ldl",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [find_nvcc()] + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
