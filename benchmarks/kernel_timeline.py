#!/usr/bin/env python
"""Where the fixed cost of a fused chain -> getAverage launch goes: globaltimer stamps inside the interpreter kernel (the library built
by `make -C finmath-lib-cuda-extensions_b200/csrc timing`, tape_interp.cuh: FMC_STAMP) for the swaption-shaped chain of
benchmarks/swaption_kernel_study.py. usage: python benchmarks/kernel_timeline.py [paths] [periods]"""
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

capi.LIB_PATH = os.path.join(ROOT, "finmath-lib-cuda-extensions_b200", "lib", "libfmcuda_timing.so")
fc.ensure_init()
L = capi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rng = np.random.default_rng(3)
base = (0.02 + 0.01 * rng.random(n)).astype(np.float32)
libor = [fc.RandomVariableCuda(0.0, base).add(0.0001 * i).add(0.0) for i in range(m)]
for v_ in libor:
    v_.getRealizationsFloat() if n <= 1 << 20 else capi.check(L.fmc_sync())
numeraire = fc.RandomVariableCuda(0.0, 1.0 + 0.1 * rng.random(n))


def chain():
    v = fc.RandomVariableCuda(0.0)
    for i in reversed(range(m)):
        li = libor[i]
        v = v.add(li.sub(0.02).mult(0.5)).discount(li, 0.5)
    return v.floor(0.0).div(numeraire).mult(1.0 / n)


NAMES = ["kernel entry (block 0)", "tape in shared memory, barriers initialised", "prologue done (TMA copies issued)", "chunk interpreted",
         "epilogue entered", "block partial stored, before the ticket", "ticket taken", "last block: starts the merge", "merge done",
         "result published"]
stamps = (ctypes.c_ulonglong * 16)()
acc = np.zeros(10)
wall = 0.0
reps = 200
for r in range(reps + 20):
    v = chain()
    t0 = time.perf_counter()
    v.getAverage()
    t1 = time.perf_counter()
    capi.check(L.fmc_debug_stamps(stamps))
    s = np.array([stamps[i] for i in range(10)], dtype=np.float64)
    if r >= 20:
        acc += s - s[0]
        wall += t1 - t0
print(f"n={n} periods={m}: getAverage() wall {wall / reps * 1e6:.1f} us (code generation + launch + kernel + result)")
prev = 0.0
for k in range(10):
    t = acc[k] / reps / 1e3
    print(f"  {t:8.2f} us  (+{t - prev:6.2f})  {NAMES[k]}")
    prev = t
