#!/usr/bin/env python
"""Three short launches for `ncu --set full`: an unfused vec+vec op, a plain getAverage and one LMM Euler time step
(the dominant kernel of the headline bench), each at a size well above L2. Usage: python benchmarks/profile_cases.py [paths]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
n = 1 << 26
rng = np.random.default_rng(1)
x = fc.RandomVariableCuda(0.0, rng.random(n, dtype=np.float32).astype(np.float64))
y = fc.RandomVariableCuda(0.0, rng.random(n, dtype=np.float32).astype(np.float64))
for _ in range(2):
    z = x.add(y)
    capi.check(capi.load().fmc_sync())          # launch: add(vec)
    z = x.add(1.5)
    capi.check(capi.load().fmc_sync())          # launch: add(scalar)
    a = x.getAverage()                          # launch: getAverage
    p = x.sub(0.5).floor(0.0).div(1.1).getAverage()   # launch: payoff chain -> getAverage
del x, y, z
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
v = m.step()          # 75 Euler-step kernels (simulation) + 144 swaption kernels (fused chain -> getAverage)
capi.check(capi.load().fmc_sync())
print("ok", a, p, v[:2])
