#!/usr/bin/env python3
"""Time mass_properties of the airfoil config (resolution 0.25, 64^3 blocks) on the specialised tier."""
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

_lib.init(0)
a = load_scenes()["cfg_airfoil"]
scene = a.compiled()
codecad_b200.mass_properties(scene, 0.25, 64)
scene.program_buffer().wait_specialized(ProgramBuffer.SINK_MASS)
for _ in range(3):
    l0, p0 = _lib.counters()
    t0 = time.perf_counter()
    mp = codecad_b200.mass_properties(scene, 0.25, 64)
    dt = (time.perf_counter() - t0) * 1e3
    l1, p1 = _lib.counters()
    print("mass_properties %.3f ms, %d launches, %d points, volume %r" % (dt, l1 - l0, p1 - p0, mp.volume))
