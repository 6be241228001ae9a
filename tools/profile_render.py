#!/usr/bin/env python3
"""Render one scene a few times with the ray caster (for ncu):  scene [width height reps tier]
   tier: 0 = interpreter kernel (default), 1 = scene-specialised kernel"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from codecad_b200 import _lib  # noqa: E402
from codecad_b200.rendering import ray_caster  # noqa: E402
from scenes import load_scenes  # noqa: E402

name = sys.argv[1]
size = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 768)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
tier = int(sys.argv[5]) if len(sys.argv) > 5 else 0
L = _lib.init(0)
_lib.check(L.cc_set_jit_mode(0))
scene = load_scenes()[name].compiled()
if tier:
    print("compile %.2f s" % scene.program_buffer().specialize(1, 16))
cam = ray_caster.get_camera_params(scene.bounding_box(), size, None)
for _ in range(reps):
    stats = {}
    t0 = time.perf_counter()
    ray_caster.render(scene, size=size, stats=stats, *cam)
    print("%s %dx%d: %.2f ms wall, kernel %.3f ms, %d evaluations (%.1f per pixel), %.2f Geval/s" % (
        name, size[0], size[1], (time.perf_counter() - t0) * 1e3, stats["ms"], stats["evaluations"],
        stats["evaluations"] / (size[0] * size[1]), stats["evaluations"] / stats["ms"] / 1e6))
