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
