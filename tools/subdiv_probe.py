#!/usr/bin/env python3
"""subdivision of a fixture scene a few times (for an ncu launch list):  scene divisor grid [reps]
(resolution = 100 / divisor: `cfg_csg_example 512 16` is the bench leg)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import codecad_b200  # noqa: E402
from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util.buffer import ProgramBuffer  # noqa: E402
from scenes import load_scenes  # noqa: E402

name, div, grid = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
_lib.init(0)
s = load_scenes()[name]
scene = s.compiled()
res = 100.0 / div   # (csg_example: 100 / 512)
print("ready %d, compile %.2f s" % scene.program_buffer().wait_specialized(ProgramBuffer.SINK_CLASSIFY))
codecad_b200.subdivision(scene, res, True, grid)
ts = []
for _ in range(reps):
    l0 = _lib.counters()[0]
    t0 = time.perf_counter()
    r = codecad_b200.subdivision(scene, res, True, grid)
    ts.append((time.perf_counter() - t0) * 1e3)
    l1 = _lib.counters()[0]
print("ms", sorted(ts), "leaf blocks", len(r[2]), "launches", l1 - l0)
