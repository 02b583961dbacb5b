#!/usr/bin/env python
"""Where does a swaption-valuation kernel spend its time? The chain of Swaption.getValue (drivers/workloads.hpp) over m swap
periods at 1 Mi paths, with the per-period operations swapped out one at a time; kernel time from the profile API."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402

fc.ensure_init()
L = capi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(3)
base = (0.02 + 0.01 * rng.random(n)).astype(np.float32)
libor = [fc.RandomVariableCuda(0.0, base).add(0.0001 * i).add(0.0) for i in range(m)]
for v_ in libor: v_.getRealizationsFloat() if n <= 1 << 20 else capi.check(L.fmc_sync())
numeraire = fc.RandomVariableCuda(0.0, 1.0 + 0.1 * rng.random(n))


def chain(kind):
    v = fc.RandomVariableCuda(0.0)
    for i in reversed(range(m)):
        li = libor[i]
        if kind == "full":            v = v.add(li.sub(0.02).mult(0.5)).discount(li, 0.5)
        elif kind == "no discount":   v = v.add(li.sub(0.02).mult(0.5)).mult(0.99)
        elif kind == "accrue instead": v = v.add(li.sub(0.02).mult(0.5)).accrue(li, 0.5)
        elif kind == "payoff only":   v = v.add(li.sub(0.02).mult(0.5))
        elif kind == "sum of leaves": v = v.add(li)
    if kind == "empty": v = libor[0].add(1.0)
    if kind == "leaf (streaming reduce kernel)": return libor[0].getAverage()
    return v.floor(0.0).div(numeraire).mult(1.0 / n).getAverage()


for kind in ("full", "no discount", "accrue instead", "payoff only", "sum of leaves", "empty", "leaf (streaming reduce kernel)"):
    for _ in range(3):
        chain(kind)
    capi.set_option("profile", 1); capi.profile_read()
    reps = 20
    for _ in range(reps):
        chain(kind)
    pr = capi.profile_read(); capi.set_option("profile", 0)
    print(f"n={n} {kind:32s} kernel {pr['tape_ms'] / reps * 1e3:7.1f} us per valuation ({pr['tape_launches'] // reps} launch)")
