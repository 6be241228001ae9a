#!/usr/bin/env python3
"""Quick kernel-time sweep on the GPU box: scene x grid size x (points/thread, program space).
Prints Gpts/s and the fraction of the FP32 roofline (static min flop/pt).  Not a bench."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from codecad_b200 import _lib  # noqa: E402
from codecad_b200.cl_util import Buffer  # noqa: E402
from codecad_b200.geometry import FLOAT4  # noqa: E402
from scenes import load_scenes  # noqa: E402


def main():
    L = _lib.init(0)
    info = _lib.device_info()
    peak = info.sm_count * 128 * 2 * info.sm_clock_khz * 1e3
    print("device", info.name.decode(), "SMs", info.sm_count, "clock kHz", info.sm_clock_khz, "fp32 peak TF", peak / 1e12)
    scenes = load_scenes()
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg_csg_example", "cfg_menger_sponge", "cfg_airfoil", "cfg_planetary", "cfg_synthetic500"]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    variants = [(4, 2), (2, 2), (1, 2), (4, 1), (2, 1), (1, 1)]
    if os.environ.get("PROBE_VARIANTS"):
        variants = [tuple(int(x) for x in v.split(":")) for v in os.environ["PROBE_VARIANTS"].split(",")]
    out = Buffer(FLOAT4, (n, n, n))
    for name in names:
        s = scenes[name]
        prog = s.compiled().program_buffer()
        pi = prog.info
        corner, step = s.grid(n)
        nx = n if name != "cfg_synthetic500" else max(8, n // 8)
        print("%s: %d micro-ops, %d words, %d slots, %d fused, flops %d..%d" % (
            name, pi.n_micro_ops, pi.n_micro_words, pi.n_slots, pi.n_fused, pi.flops_min, pi.flops_max))
        for pts, space in (variants if os.environ.get("PROBE_INTERP", "1") != "0" else []):
            _lib.check(L.cc_set_tuning(pts, space))
            best = None
            for it in range(3):
                e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
                _lib.check(L.cc_event_record(ctypes.byref(e0)))
                rc = L.cc_grid_eval(prog.handle, _lib.f3(corner), float(step), nx, n, n, 0, 0, out.device_ptr, None)
                if rc < 0:
                    print("   pts=%d space=%d: %s" % (pts, space, L.cc_last_error().decode()))
                    break
                _lib.check(L.cc_event_record(ctypes.byref(e1)))
                _lib.check(L.cc_event_wait(e1))
                ms = ctypes.c_float()
                _lib.check(L.cc_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
                L.cc_event_destroy(e0)
                L.cc_event_destroy(e1)
                if it and (best is None or ms.value < best):
                    best = ms.value
            if best is None:
                continue
            pts_s = nx * n * n / (best * 1e-3)
            print("   pts=%d space=%s: %8.3f ms  %8.3f Gpts/s  fp32 frac(min flops) %.3f  store GB/s %.0f" % (
                pts, {1: "const", 2: "smem", 3: "hybrid"}[space], best, pts_s / 1e9, pts_s * pi.flops_min / peak, pts_s * 16 / 1e9))
        # scene-specialised kernel (NVRTC)
        if os.environ.get("PROBE_JIT", "1") != "0":
            for jp in [int(x) for x in os.environ.get("PROBE_JIT_PTS", "2,1").split(",")]:
                _lib.check(L.cc_set_tuning(0, 0))
                try:
                    secs = prog.specialize(jp, 1)
                except Exception as exc:  # noqa: BLE001
                    print("   jit pts=%d: %s" % (jp, str(exc)[:300]))
                    break
                best = None
                for it in range(3):
                    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
                    _lib.check(L.cc_event_record(ctypes.byref(e0)))
                    _lib.check(L.cc_grid_eval(prog.handle, _lib.f3(corner), float(step), nx, n, n, 0, 0, out.device_ptr, None))
                    _lib.check(L.cc_event_record(ctypes.byref(e1)))
                    _lib.check(L.cc_event_wait(e1))
                    ms = ctypes.c_float()
                    _lib.check(L.cc_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
                    if it and (best is None or ms.value < best):
                        best = ms.value
                pts_s = nx * n * n / (best * 1e-3)
                import zlib
                head = np.empty((2, n, n), dtype=FLOAT4)
                _lib.check(L.cc_memcpy_d2h_async(head.ctypes.data, out.device_ptr, head.nbytes, None))
                _lib.check(L.cc_synchronize())
                print("   JIT pts=%d (compile %.1f s): %8.3f ms  %8.3f Gpts/s  fp32 frac(min flops) %.3f  crc %08x" % (
                    jp, secs, best, pts_s / 1e9, pts_s * pi.flops_min / peak, zlib.crc32(head.tobytes())))
            prog.use_specialized(False)
    _lib.check(L.cc_set_tuning(0, 0))


if __name__ == "__main__":
    main()
