#!/usr/bin/env python
"""Small-vector latency of the boundary: what one flush / one reduction costs end to end when the vector is tiny and
the GPU work is negligible (the regime of the reference's 5k-50k path calibrations, README.md:24-28)."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402

fc.ensure_init()
L = capi.load()


def host_prof():
    out = {}
    for key in ("host_us_codegen", "host_us_launch", "host_us_sync"):
        v = ctypes.c_double()
        capi.check(L.fmc_get_option(key.encode(), ctypes.byref(v)))
        out[key] = v.value
    return out


for n in (4096, 65536, 1 << 20):
    x = fc.RandomVariableCuda(0.0, np.random.rand(n))
    y = fc.RandomVariableCuda(0.0, np.random.rand(n))
    reps = 300
    for name, fn in (("add+sync", lambda: (x.add(y), capi.check(L.fmc_sync()))),
                     ("getAverage(leaf)", lambda: x.getAverage()),
                     ("10-op chain -> getAverage", lambda: x.mult(y).add(1.0).sub(y).mult(0.5).add(x).floor(0.0).div(1.1).sub(x).abs().add(y).getAverage())):
        for _ in range(20):
            fn()
        capi.check(L.fmc_reset_stats())
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps * 1e6
        hp = host_prof()
        print(f"n={n:8d} {name:28s} {dt:8.1f} us/call   host: codegen {hp['host_us_codegen'] / reps:6.1f}  launch {hp['host_us_launch'] / reps:6.1f}  "
              f"copy+sync {hp['host_us_sync'] / reps:6.1f} us")
