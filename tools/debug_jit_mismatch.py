#!/usr/bin/env python3
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from scenes import load_scenes
import codecad_b200
from codecad_b200 import _lib
from codecad_b200.cl_util.buffer import ProgramBuffer
_lib.init(0)
_lib.lib().cc_set_jit_mode(0)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg_planetary"
s = load_scenes()[name]
dims = (16, 9, 37)
corner, step = s.grid(40)
want = oracle.grid_eval(s.words, corner, step, dims)
f4 = lambda a: np.stack([a["x"], a["y"], a["z"], a["w"]], axis=-1)
for T, MB, pts in [(128, 0, 2), (512, 2, 2), (512, 0, 2), (256, 4, 2), (512, 2, 1), (256, 2, 4)]:
    os.environ["CODECAD_B200_JIT_THREADS"] = str(T)
    if MB: os.environ["CODECAD_B200_JIT_MINB"] = str(MB)
    else: os.environ.pop("CODECAD_B200_JIT_MINB", None)
    prog = ProgramBuffer(s.words)
    prog.specialize(pts, 1)
    got = f4(codecad_b200.grid_eval(prog, corner, step, dims))
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    idx = np.argwhere(bad.any(axis=-1))
    print("T=%d minb=%d pts=%d: %d mismatching points of %d" % (T, MB, pts, len(idx), want[..., 0].size))
    for i in idx[:6]:
        lin = (i[0] * dims[1] + i[1]) * dims[2] + i[2]
        print("   ", tuple(i), "linear", lin, "got", got[tuple(i)], "want", want[tuple(i)])
