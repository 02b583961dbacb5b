#!/usr/bin/env python
"""Host-side breakdown of the end-to-end LMM step (Brownian increments uploaded from host doubles every step).
usage: FMC_OPTIONS=... FMC_HOST_THREADS=k python benchmarks/e2e_phases.py [paths]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
m.prepare_host_brownian()
L = capi.load()
KEYS = ("host_us_codegen", "host_us_launch", "host_us_sync", "host_us_upload", "host_us_upload_wait")


def host():
    out = []
    for k in KEYS:
        v = ctypes.c_double(); capi.check(L.fmc_get_option(k.encode(), ctypes.byref(v))); out.append(v.value)
    return out


for from_host in (False, True):
    for _ in range(2):
        m.step(None, from_host=from_host)
    h0 = host(); t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        m.step(None, from_host=from_host)
    wall = (time.perf_counter() - t0) / reps * 1e3
    h = [(b - a) / reps / 1e3 for a, b in zip(h0, host())]
    print(f"paths={paths} from_host={from_host}: wall {wall:6.2f} ms  host: codegen {h[0]:5.2f} launch {h[1]:5.2f} sync {h[2]:5.2f} upload {h[3]:5.2f} (waiting for a staging chunk {h[4]:5.2f})")
