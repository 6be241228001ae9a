#!/usr/bin/env python3
"""mass_properties of a fixture scene a few times (for an ncu launch list):  scene resolution grid [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import codecad_b200  # noqa: E402
from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util.buffer import ProgramBuffer  # noqa: E402
from scenes import load_scenes  # noqa: E402

name, res, grid = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
_lib.init(0)
scene = load_scenes()[name].compiled()
print("ready %d, compile %.2f s" % scene.program_buffer().wait_specialized(ProgramBuffer.SINK_MASS))
stats = {}
codecad_b200.mass_properties(scene, res, grid, stats=stats)
print(stats)
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    r = codecad_b200.mass_properties(scene, res, grid)
    ts.append((time.perf_counter() - t0) * 1e3)
print("ms", sorted(ts), "volume", r.volume)
