#!/usr/bin/env python3
"""The e2e leg of bench.py alone: planetary 1024^3 through cc_grid_eval_to_host into pinned host memory
(fresh program per step), for the slab size in $CODECAD_B200_SLAB_MIB:  [n] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util.buffer import ProgramBuffer, _Pinned  # noqa: E402
from codecad_b200.geometry import FLOAT4  # noqa: E402
from scenes import load_scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.init(0)
s = load_scenes()["cfg_planetary"]
corner, step = s.grid(n)
c3 = _lib.f3(corner)
prog = ProgramBuffer(s.words)
print("ready %d, compile %.2f s" % prog.wait_specialized(ProgramBuffer.SINK_FLOAT4))
pin = _Pinned(n * n * n * 16)
host = pin.array(FLOAT4, (n, n, n))


def once():
    p = ProgramBuffer(s.words)
    _lib.check(L.cc_grid_eval_to_host(p.handle, c3, float(step), n, n, n, 0, 0, host.ctypes.data))
    p.release()


once()
ts = []
for _ in range(steps):
    t0 = time.perf_counter()
    once()
    ts.append(time.perf_counter() - t0)
best = min(ts)
print("slab %s MiB: %.1f ms per grid (best of %d), %.2f Gpts/s, %.1f GB/s" % (os.environ.get("CODECAD_B200_SLAB_MIB", "default"), best * 1e3, steps,
                                                                          n ** 3 / best / 1e9, n ** 3 * 16 / best / 1e9))
