#!/usr/bin/env python3
"""Pinned-memory D2H / H2D bandwidth of the box (what bounds the e2e leg of bench.py)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codecad_b200 import _lib
L = _lib.init(0)
for mb in (64, 256, 1024, 4096):
    n = mb << 20
    d, h = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.cc_buffer_alloc(n, ctypes.byref(d)))
    _lib.check(L.cc_host_alloc(n, ctypes.byref(h)))
    ctypes.memset(h, 1, n)
    for name, fn in (("D2H", lambda: L.cc_memcpy_d2h_async(h, d, n, None)), ("H2D", lambda: L.cc_memcpy_h2d_async(d, h, n, None))):
        fn(); L.cc_synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        L.cc_synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("%s %5d MiB  %.1f GB/s" % (name, mb, n / dt / 1e9))
    L.cc_buffer_free(d); L.cc_host_free(h)

# the e2e pattern: a 16 GiB pinned buffer filled by 64 MiB chunk copies from a small device ring
n = 16 << 30
chunk = 64 << 20
d, h = ctypes.c_void_p(), ctypes.c_void_p()
_lib.check(L.cc_buffer_alloc(chunk * 4, ctypes.byref(d)))
t0 = time.perf_counter()
_lib.check(L.cc_host_alloc(n, ctypes.byref(h)))
print("cc_host_alloc 16 GiB: %.2f s" % (time.perf_counter() - t0))
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(n // chunk):
        L.cc_memcpy_d2h_async(ctypes.c_void_p(h.value + i * chunk), ctypes.c_void_p(d.value + (i % 4) * chunk), chunk, None)
    L.cc_synchronize()
    dt = time.perf_counter() - t0
    print("D2H 16 GiB in 64 MiB chunks: %.1f GB/s (%.0f ms)" % (n / dt / 1e9, dt * 1e3))
import subprocess
print(subprocess.run("numactl -H 2>/dev/null | head -8; nvidia-smi topo -m 2>/dev/null | head -6; lscpu | grep -i 'numa\\|socket\\|model name'", shell=True, capture_output=True, text=True).stdout)
