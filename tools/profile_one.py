#!/usr/bin/env python3
"""Launch the grid_eval kernel a few times for one scene (for ncu):
   scene n pts space [reps] [jit_pts]      (jit_pts > 0: scene-specialised kernel)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util import Buffer  # noqa: E402
from codecad_b200.geometry import FLOAT4  # noqa: E402
from scenes import load_scenes  # noqa: E402

name, n, pts, space = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
jit = int(sys.argv[6]) if len(sys.argv) > 6 else 0
L = _lib.init(0)
s = load_scenes()[name]
prog = s.compiled().program_buffer()
_lib.check(L.cc_set_jit_mode(0))   # profile exactly the tier asked for
if jit == 9:   # what a user gets in steady state: every specialised kernel of dense float4 grids, part culling included
    from codecad_b200.cl_util.buffer import ProgramBuffer
    _lib.check(L.cc_set_jit_mode(1))
    print("ready %d, compile %.2f s" % prog.wait_specialized(ProgramBuffer.SINK_FLOAT4))
elif jit:
    print("compile %.2f s" % prog.specialize(jit, 1))
corner, step = s.grid(n)
out = Buffer(FLOAT4, (n, n, n))
_lib.check(L.cc_set_tuning(pts, space))
import time  # noqa: E402
_lib.check(L.cc_grid_eval(prog.handle, _lib.f3(corner), float(step), n, n, n, 0, 0, out.device_ptr, None))
_lib.check(L.cc_synchronize())
t0 = time.perf_counter()
for _ in range(reps):
    _lib.check(L.cc_grid_eval(prog.handle, _lib.f3(corner), float(step), n, n, n, 0, 0, out.device_ptr, None))
_lib.check(L.cc_synchronize())
ms = (time.perf_counter() - t0) * 1e3 / reps
print("ok  %.3f ms per grid (host clock around %d launches), %.2f Gpts/s" % (ms, reps, n ** 3 / ms / 1e6))
