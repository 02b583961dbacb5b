#!/usr/bin/env python
"""Windows of several time steps per launch (option window_levels) against one launch per step: the LIBORs of an LMM simulation
and the 144 swaption values must be bit-identical (same arithmetic per path, different schedule).
usage: python benchmarks/window_check.py [paths]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
D = DriverLib()
probes = [(1, 5), (2, 2), (7, 7), (7, 8), (8, 79), (33, 34), (40, 60), (41, 41), (79, 79), (80, 79)]


def run(w, elems=0, threshold=4096, reduce_min=2048):
    fc.set_option("window_levels", w)
    fc.set_option("tape_elems", elems)
    fc.set_option("flush_threshold", threshold)
    fc.set_option("window_reduce_min", reduce_min)
    m = D.lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
    vals = np.asarray(m.step()).copy()
    libors = [m.libor(t, i).copy() for (t, i) in probes]
    m.close()
    return vals, libors


ref_vals, ref_libors = run(0)
ok = True
for w, e in ((2, 0), (3, 0), (4, 0), (3, 8), (5, 8), (2, 4)):
    vals, libors = run(w, e)
    # (another chunk geometry sums the paths of a swaption value in another order: last-bit differences there)
    same_v = np.array_equal(vals, ref_vals) if e == 0 else bool(np.allclose(vals, ref_vals, rtol=1e-12, atol=0))
    same_l = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(libors, ref_libors))
    print(f"window_levels={w} tape_elems={e}: swaption values identical {same_v}, LIBORs bit-identical {same_l}")
    ok = ok and same_v and same_l
# no automatic flush at all: the whole simulation is still pending when the first swaption is valued and goes through the windows there
vals, libors = run(3, 0, threshold=10_000_000, reduce_min=0)
# (a valuation that reads a differently flushed simulation may pick another chunk geometry for its sum: last-bit differences)
same_v = bool(np.allclose(vals, ref_vals, rtol=1e-12, atol=0))
same_l = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(libors, ref_libors))
print(f"window_levels=3, simulation flushed by the first valuation: swaption values equal to 1e-12 {same_v} (max rel diff {float(np.max(np.abs(vals - ref_vals) / np.abs(ref_vals))):.3g}), LIBORs bit-identical {same_l}")
ok = ok and same_v and same_l
fc.set_option("flush_threshold", 4096); fc.set_option("window_reduce_min", 2048)


# a three-factor model shares three running sums and three Brownian increments per time step between its components: windows of
# three levels do not fit the register file, the window size adapts (Runtime::run_windows) — same LIBORs, no more launches than
# one per time step plus the few windows it took to find out
def run3(w):
    fc.set_option("window_levels", w)
    fc.set_option("tape_elems", 0)
    m = D.lmm(paths, 40, 0.5, 3, 31415, 0, (0, paths))
    for _ in range(2):
        m.simulate(); fc.sync()
    k0 = fc.stats()["n_kernels"]
    m.simulate(); fc.sync()
    launches = fc.stats()["n_kernels"] - k0
    libors = [m.libor(t, i).copy() for (t, i) in ((1, 5), (7, 8), (20, 30), (39, 39), (40, 39))]
    m.close()
    return launches, libors


l0, ref3 = run3(0)
l3, got3 = run3(3)
same_l = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(got3, ref3))
print(f"three-factor LMM: LIBORs bit-identical {same_l}; launches per simulation {l3} with adaptive windows, {l0} with one launch per time step")
ok = ok and same_l and l3 <= l0 + 8
fc.set_option("window_levels", 3); fc.set_option("tape_elems", 0)
sys.exit(0 if ok else 1)
