#!/usr/bin/env python
"""Brownian generation rate (MT19937 + AS241 on the device), 1 Mi paths x 80 steps by default; the start states of a seed are cached
after its first generation (the cold jump-ahead is reported separately).  usage: python benchmarks/brownian_rate.py [paths] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 80
td = fc.TimeDiscretization(0.0, steps, 0.5)
ts = []
for i in range(8):
    capi.check(capi.load().fmc_sync()); t0 = time.perf_counter()
    bm = fc.BrownianMotionCuda(td, 1, paths, 4711)
    inc = bm.getBrownianIncrement(0, 0); capi.check(capi.load().fmc_sync()); ts.append(time.perf_counter() - t0); del bm, inc
print(f"paths {paths} steps {steps}: first {1e3 * ts[0]:.2f} ms, cached start states {1e3 * min(ts[1:]):.3f} ms = {paths * steps / min(ts[1:]) / 1e9:.1f} G increments/s")
