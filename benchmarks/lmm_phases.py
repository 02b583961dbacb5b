#!/usr/bin/env python
"""Where one LMM calibration step spends its wall time: the simulation phase (asynchronous launches, one sync at the end) versus
the swaption phase (one fused reduction and one host read per product), with the host-side shares of each.
usage: FMC_OPTIONS=... python benchmarks/lmm_phases.py [paths]"""
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
import ctypes  # noqa: E402

KEYS = ("host_us_codegen", "host_us_launch", "host_us_sync")


def host():
    out = []
    for k in KEYS:
        v = ctypes.c_double()
        capi.check(L.fmc_get_option(k.encode(), ctypes.byref(v)))
        out.append(v.value)
    return out


def timed(fn, reps=5):
    for _ in range(2):
        fn(); capi.check(L.fmc_sync())
    h0 = host(); t0 = time.perf_counter()
    for _ in range(reps):
        fn(); capi.check(L.fmc_sync())
    wall = (time.perf_counter() - t0) / reps * 1e3
    h = [(b - a) / reps / 1e3 for a, b in zip(h0, host())]
    return wall, h


def report(ws, hs, wf, hf):
    print(f"paths={paths}  simulate: wall {ws:6.2f} ms (host codegen {hs[0]:5.2f} launch {hs[1]:5.2f} sync {hs[2]:5.2f})")
    print(f"paths={paths}  full step: wall {wf:6.2f} ms (host codegen {hf[0]:5.2f} launch {hf[1]:5.2f} sync {hf[2]:5.2f})")
    print(f"paths={paths}  swaption phase = {wf - ws:6.2f} ms for {len(m.products())} products = {(wf - ws) / len(m.products()) * 1e3:6.1f} us each")


for prof in (0, 1):
    capi.set_option("profile", prof)
    capi.profile_read()
    ws, hs = timed(m.simulate)
    ps = capi.profile_read()
    wf, hf = timed(m.step)
    pf = capi.profile_read()
    print(f"-- per-launch event timing {'on' if prof else 'off'}")
    report(ws, hs, wf, hf)
    if prof:
        reps = 7      # 2 warm-up + 5 timed calls of each
        ks, kf = ps["tape_ms"] / reps, pf["tape_ms"] / reps
        bs, bf = ps["tape_algorithmic_bytes"] / reps, pf["tape_algorithmic_bytes"] / reps
        print(f"paths={paths}  kernels: simulation {ks:6.3f} ms ({ps['tape_launches'] / reps:.0f} launches, {bs / 1e9:6.2f} GB algorithmic, {bs / ks / 1e6:7.1f} GB/s)")
        print(f"paths={paths}  kernels: swaptions  {kf - ks:6.3f} ms ({(pf['tape_launches'] - ps['tape_launches']) / reps:.0f} launches, {(bf - bs) / 1e9:6.2f} GB algorithmic, "
              f"{(bf - bs) / max(kf - ks, 1e-9) / 1e6:7.1f} GB/s)")
        print(f"paths={paths}  kernels: whole step {kf:6.3f} ms, {bf / 1e9:6.2f} GB algorithmic, {bf / kf / 1e6:7.1f} GB/s")


