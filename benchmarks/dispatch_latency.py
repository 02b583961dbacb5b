#!/usr/bin/env python
"""Per-instruction latency of the interpreter with ONE warp per SM sub-partition (n = SMs * 4 * 512 elements, one chunk per
warp): long chains of a single instruction kind, kernel time / chain length = cycles per dispatch for a lone warp."""
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
warps_per_sched = int(sys.argv[1]) if len(sys.argv) > 1 else 1
elems = int(sys.argv[2]) if len(sys.argv) > 2 else 16
capi.set_option("tape_elems", elems)
n = 148 * 4 * 32 * elems * warps_per_sched
K = 600
x = fc.RandomVariableCuda(0.0, np.random.rand(n) + 0.5)
y = fc.RandomVariableCuda(0.0, np.random.rand(n) + 0.5)


def chain(kind):
    c = x
    for k in range(K):
        if kind == "MUL_I": c = c.mult(1.0001)
        elif kind == "ADD_I": c = c.add(0.001)
        elif kind == "SQUARED": c = c.squared().add(0.5) if k % 2 else c.abs()
        elif kind == "ADD_S(leaf)": c = c.add(y)
        elif kind == "DIV_I": c = c.div(1.0001)
        elif kind == "VID_I": c = c.vid(1.0001)
        elif kind == "MULADD_II": c = c.mult(1.0001).add(0.001)
        elif kind == "DISCOUNT_S": c = c.discount(y, 0.001)
        elif kind == "MIX of 8 cheap handlers":
            c = (c.mult(1.0001), c.add(0.001), c.sub(0.001), c.bus(3.0), c.cap(1e9), c.floor(-1e9), c.abs(), c.squared())[k % 8] if k % 8 != 7 else c.sqrt()
    return c


capi.set_option("flush_threshold", 1e9)
capi.set_option("fuse_ops", 0)
for kind in ("MUL_I", "ADD_I", "MIX of 8 cheap handlers", "ADD_S(leaf)", "DIV_I", "VID_I", "DISCOUNT_S"):
    for _ in range(2):
        r = chain(kind); capi.check(L.fmc_sync()); del r
    capi.set_option("profile", 1); capi.profile_read()
    reps = 3
    for _ in range(reps):
        r = chain(kind); capi.check(L.fmc_sync()); del r
    pr = capi.profile_read(); capi.set_option("profile", 0)
    st = fc.stats()
    us = pr["tape_ms"] / reps * 1e3
    per = us / K * 1e3
    print(f"{kind:26s} elems={elems} warps/scheduler={warps_per_sched}  kernel {us:8.1f} us for {K} ops -> {per:6.1f} ns = {per * 1.965:6.0f} cycles per op (at 1965 MHz), launches/rep {pr['tape_launches'] // reps}")
