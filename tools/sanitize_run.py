#!/usr/bin/env python3
"""Small launches of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
ordered compaction with decoupled look-back (CLASSIFY, MASS), the exact mass accumulator, marching
cubes (count / scan / emit), the union-forest kernel (both launches), outlines, the ray caster.
Interpreter tier only unless --jit (NVRTC kernels are checked the same way, but compile first).
    compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import codecad_b200  # noqa: E402
from codecad_b200 import _lib  # noqa: E402
from codecad_b200.rendering import image, mesh_arrays, polygon2d  # noqa: E402
from scenes import load_scenes  # noqa: E402

L = _lib.init(0)
jit = "--jit" in sys.argv
_lib.check(L.cc_set_jit_mode(2 if jit else 0))
S = load_scenes()
box = S["sub_box10"].compiled()
print("subdivision", len(codecad_b200.subdivision(box, 1, True, 4)[2]))
csg = S["cfg_csg_example"].compiled()
print("subdivision csg", len(codecad_b200.subdivision(csg, 100 / 64, True, 8)[2]))
print("mass", codecad_b200.mass_properties(S["mp_drunk_box"].compiled(), 0.2, 8).volume)
print("mass airfoil", codecad_b200.mass_properties(S["cfg_airfoil"].compiled(), 4.0, 16).volume)
v, b, boxes = mesh_arrays(codecad_b200.CompiledScene(S["sub_box10"].words, 3, S["sub_box10"].box_a, S["sub_box10"].box_b, 1.0, "box"), 8)
print("mesh", len(v), len(boxes))
for name in ("cfg_synthetic32", "forest_dense64"):
    s = S[name]
    corner, step = s.grid(40)
    g = codecad_b200.grid_eval(s.compiled(), corner, step, (40, 37, 33))
    print("forest", name, float(np.nanmin(g["w"])))
gear = S["dsdf2d_gear"].compiled()
print("outlines", len(list(polygon2d.polygon(gear, 16))))
t = S["dsdf3d_csg_thing"].compiled()
print("picture", image.render_pixels(t, (48, 32)).shape)
_lib.check(L.cc_synchronize())
print("ok")
