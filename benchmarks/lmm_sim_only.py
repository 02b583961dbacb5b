#!/usr/bin/env python
"""Kernel time of the LMM Euler simulation alone (no swaption valuation): tuning aid for the interpreter's scheduling knobs.
usage: FMC_OPTIONS=... python benchmarks/lmm_sim_only.py [paths]"""
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
L = capi.load()
for _ in range(2):
    m.simulate(); capi.check(L.fmc_sync())
capi.set_option("profile", 1); capi.profile_read()
reps = 4
t0 = time.perf_counter()
for _ in range(reps):
    m.simulate(); capi.check(L.fmc_sync())
wall = (time.perf_counter() - t0) / reps * 1e3
pr = capi.profile_read()
print(f"FMC_OPTIONS={os.environ.get('FMC_OPTIONS', '')!r:45s} paths={paths} sim wall {wall:7.2f} ms  kernels {pr['tape_ms'] / reps:7.2f} ms  launches {pr['tape_launches'] // reps}  "
      f"alg GB/s {pr['tape_algorithmic_bytes'] / (pr['tape_ms'] * 1e-3) / 1e9:7.1f}")
