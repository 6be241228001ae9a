#!/usr/bin/env python3
"""mesh_arrays of a config scene, a few times (for ncu):  [scene [feature_size [grid]]]
(default: csg_example at 512^3 effective resolution, 128^3 blocks)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codecad_b200 import CompiledScene, _lib  # noqa: E402
from codecad_b200.cl_util.buffer import ProgramBuffer  # noqa: E402
from codecad_b200.rendering import mesh_arrays  # noqa: E402
from scenes import load_scenes  # noqa: E402

_lib.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg_csg_example"
feature = float(sys.argv[2]) if len(sys.argv) > 2 else 2 * 100.0 / 512
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 128
c = load_scenes()[name]
scene = CompiledScene(c.words, 3, c.box_a, c.box_b, feature, name)
mesh_arrays(scene, grid)
scene.program_buffer().wait_specialized(ProgramBuffer.SINK_PYMCUBES | ProgramBuffer.SINK_CLASSIFY)
for _ in range(3):
    t0 = time.perf_counter()
    vertices, block, boxes = mesh_arrays(scene, grid)
    print("mesh_arrays %.3f ms, %d triangles, %d blocks" % ((time.perf_counter() - t0) * 1e3, len(vertices), len(boxes)))
