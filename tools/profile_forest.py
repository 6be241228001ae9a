#!/usr/bin/env python3
"""Launch the union-forest kernel on a slab of the 2048^3 grid of the 500-box scene (for ncu):
   [planes] [reps] [scene]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util import Buffer  # noqa: E402
from codecad_b200.geometry import FLOAT4  # noqa: E402
from scenes import load_scenes  # noqa: E402

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
name = sys.argv[3] if len(sys.argv) > 3 else "cfg_synthetic500"
n = 2048
L = _lib.init(0)
s = load_scenes()[name]
prog = s.compiled().program_buffer()
corner, step = s.grid(n)
out = Buffer(FLOAT4, (planes, n, n))
for r in range(reps):
    _lib.check(L.cc_synchronize())
    t0 = time.perf_counter()
    _lib.check(L.cc_grid_eval(prog.handle, _lib.f3(corner), float(step), planes, n, n, 1024 - planes // 2, 0, out.device_ptr, None))
    _lib.check(L.cc_synchronize())
    dt = time.perf_counter() - t0
    print("rep %d: %.3f ms  %.1f Gpts/s" % (r, dt * 1e3, planes * n * n / dt / 1e9))
print("ok")
