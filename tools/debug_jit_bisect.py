#!/usr/bin/env python3
"""Bisect the first micro-op at which the packed (pts=2) specialised kernel differs from the scalar (pts=1) one."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
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
f4 = lambda a: np.stack([a["x"], a["y"], a["z"], a["w"]], axis=-1)
os.environ["CODECAD_B200_JIT_THREADS"] = "128"
os.environ["CODECAD_B200_JIT_MINB"] = "0"
def run(pts, stop, nopack=False):
    os.environ["CODECAD_B200_JIT_STOP"] = str(stop)
    if nopack: os.environ["CODECAD_B200_JIT_NOPACK"] = "1"
    else: os.environ.pop("CODECAD_B200_JIT_NOPACK", None)
    prog = ProgramBuffer(s.words)
    prog.specialize(pts, 1)
    return f4(codecad_b200.grid_eval(prog, corner, step, dims))
def nbad(a, b):
    return int((~((a == b) | (np.isnan(a) & np.isnan(b)))).any(axis=-1).sum())
n_ops = ProgramBuffer(s.words).info.n_micro_ops
full1, full2 = run(1, 10**6), run(2, 10**6)
print("full: packed vs scalar mismatches", nbad(full1, full2), " nopack pts=2 vs scalar", nbad(full1, run(2, 10**6, True)))
lo = None
for k in range(int(sys.argv[2]), int(sys.argv[3])):
    b = nbad(run(1, k), run(2, k))
    print("stop after %d ops: %d mismatches" % (k, b))
    if b and lo is None:
        lo = k
        break
mc = ProgramBuffer(s.words).microcode()
pc = 0
for i in range(lo + 1):
    h = int(mc[pc]); ln = (h >> 26) * 4
    if i >= lo - 3: print("op#%d pc %d: mop %d src %d dst %d len %d params %s" % (i, pc, h & 255, (h >> 8) & 511, (h >> 17) & 511, ln, mc[pc+1:pc+ln].view(np.float32)[:8]))
    pc += ln
a, b = run(1, lo), run(2, lo)
bad = (~((a == b) | (np.isnan(a) & np.isnan(b)))).any(axis=-1)
for i in np.argwhere(bad)[:5]:
    print(tuple(int(x) for x in i), "scalar", a[tuple(i)], [hex(v) for v in a[tuple(i)].view(np.uint32)], "packed", b[tuple(i)], [hex(v) for v in b[tuple(i)].view(np.uint32)])
a0, b0 = run(1, lo - 1), run(2, lo - 1)
for i in np.argwhere(bad)[:5]:
    print("  input", tuple(int(x) for x in i), a0[tuple(i)], [hex(v) for v in a0[tuple(i)].view(np.uint32)])
