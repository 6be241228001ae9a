#!/usr/bin/env python3
"""polygon() of the gear at feature_size/16 with 32x32 boxes, a few times (for ncu)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codecad_b200 import CompiledScene, _lib  # noqa: E402
from codecad_b200.cl_util.buffer import ProgramBuffer  # noqa: E402
from codecad_b200.rendering import polygon2d  # noqa: E402
from scenes import load_scenes  # noqa: E402

_lib.init(0)
s = load_scenes()["dsdf2d_gear"]
div = int(sys.argv[1]) if len(sys.argv) > 1 else 16
scene = CompiledScene(s.words, 2, s.box_a, s.box_b, s.feature_size / div, "gear")
list(polygon2d.polygon(scene, 32))
scene.program_buffer().wait_specialized(ProgramBuffer.SINK_FLOAT4 | ProgramBuffer.SINK_CLASSIFY)
for _ in range(3):
    l0, p0 = _lib.counters()
    t0 = time.perf_counter()
    out = list(polygon2d.polygon(scene, 32))
    dt = (time.perf_counter() - t0) * 1e3
    l1, p1 = _lib.counters()
    print("polygon %.3f ms, %d launches, %d evaluations, %d outlines, %d vertices" % (
        dt, l1 - l0, p1 - p0, len(out), sum(len(c) for c in out)))
