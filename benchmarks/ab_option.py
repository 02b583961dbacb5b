#!/usr/bin/env python
"""A/B of one runtime option on the LMM step, interleaved in one process (run-to-run noise between processes is several per cent):
usage: python benchmarks/ab_option.py <option> <value A> <value B> [paths] [rounds]
prints, per value, the mean and minimum over the rounds of: wall ms per step, kernel ms per step (per-launch CUDA events)."""
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

opt, va, vb = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
paths = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 6
fc.ensure_init()
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
L = capi.load()
res = {va: {"wall": [], "kern": [], "sim": []}, vb: {"wall": [], "kern": [], "sim": []}}
for r in range(rounds + 1):
    for v in (va, vb):
        capi.set_option(opt, v)
        for _ in range(2):
            m.step()
        capi.check(L.fmc_sync())
        capi.set_option("profile", 0)
        t0 = time.perf_counter()
        for _ in range(5):
            m.step()
        capi.check(L.fmc_sync())
        wall = (time.perf_counter() - t0) / 5 * 1e3
        capi.set_option("profile", 1)
        capi.profile_read()
        for _ in range(3):
            m.simulate(); capi.check(L.fmc_sync())
        ps = capi.profile_read()
        for _ in range(3):
            m.step()
        capi.check(L.fmc_sync())
        pf = capi.profile_read()
        capi.set_option("profile", 0)
        if r > 0:
            res[v]["wall"].append(wall); res[v]["kern"].append(pf["tape_ms"] / 3); res[v]["sim"].append(ps["tape_ms"] / 3)
for v in (va, vb):
    d = res[v]
    print(f"{opt}={v:g} paths={paths}: wall mean {sum(d['wall']) / len(d['wall']):7.3f} min {min(d['wall']):7.3f} | step kernels mean {sum(d['kern']) / len(d['kern']):7.3f} min {min(d['kern']):7.3f}"
          f" | simulate-only kernels mean {sum(d['sim']) / len(d['sim']):7.3f} min {min(d['sim']):7.3f}")
