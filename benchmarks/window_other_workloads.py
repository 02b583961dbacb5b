#!/usr/bin/env python
"""Windows of time steps on workloads other than the one-factor LMM: Black-Scholes Euler Monte-Carlo (one chain, 100 steps) and a
three-factor LMM (three Brownian increments and three running sums per time step) — same values, wall time with and without windows.
usage: python benchmarks/window_other_workloads.py [paths]"""
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
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
L = capi.load()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
D = DriverLib()


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    capi.check(L.fmc_sync())
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    capi.check(L.fmc_sync())
    return (time.perf_counter() - t0) / reps * 1e3, out


res = {}
for w in (0, 3):
    fc.set_option("window_levels", w)
    ms, v = timed(lambda: D.bs_call(paths))
    m3 = D.lmm(paths, 40, 0.5, 3, 31415, 0, (0, paths))
    ms3, _ = timed(lambda: (m3.simulate(), None)[1])
    probe = m3.libor(40, 39).copy()
    vals3 = np.asarray(m3.step()).copy()
    m3.close()
    res[w] = (v, probe, vals3)
    print(f"window_levels={w}: Black-Scholes {paths} paths x 100 steps {ms:7.2f} ms (value {v[0]:.8f});  3-factor LMM 40 x 40 simulation {ms3:7.2f} ms")
same_bs = res[0][0] == res[3][0]
same_libor = np.array_equal(res[0][1].view(np.uint32), res[3][1].view(np.uint32))
# (a valuation whose inputs one mode materialised and the other fused sums its paths in another order: last-bit differences)
same_vals = bool(np.allclose(res[0][2], res[3][2], rtol=1e-12, atol=0))
print(f"Black-Scholes value identical {same_bs}; 3-factor LIBOR bit-identical {same_libor} (max abs diff {np.max(np.abs(res[0][1] - res[3][1])):.3g}); "
      f"3-factor swaption values equal to 1e-12 {same_vals} (max rel diff {np.max(np.abs(res[0][2] - res[3][2]) / np.maximum(np.abs(res[0][2]), 1e-300)):.3g})")
same = same_bs and same_libor and same_vals
fc.set_option("window_levels", 3)
sys.exit(0 if same else 1)
