#!/usr/bin/env python
"""One LMM calibration step for `ncu`: 80 Euler-step launches of the interpreter (simulation) followed by 144 fused
chain -> getAverage launches (swaptions). Usage: python benchmarks/profile_lmm.py [paths] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
for _ in range(steps):
    v = m.step()
capi.check(capi.load().fmc_sync())
print("ok", v[:2])
