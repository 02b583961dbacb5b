#!/usr/bin/env python
"""BASELINE.json configs 1 and 3 as wall-clock timings on one GPU next to the CPU restatement (single thread):
 1: Black-Scholes European call, Euler Monte Carlo, 100 steps, 100k and 1M paths (MonteCarloBlackScholesModelTest constants)
 3: Bermudan swaption under the LMM with conditional-expectation regression + choose(), 1M paths (CPU: bounded sample)."""
import json
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
from oracle.workloads_oracle import driver  # noqa: E402

fc.ensure_init()
gpu, cpu = DriverLib(), driver()
L = capi.load()
out = {}


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        capi.check(L.fmc_sync())
        t0 = time.perf_counter(); r = fn(); capi.check(L.fmc_sync()); ts.append(time.perf_counter() - t0)
    return min(ts), r


for paths in (100_000, 1_000_000):
    tg, (vg, an) = timed(lambda: gpu.bs_call(paths))
    row = {"gpu_s": tg, "gpu_value": vg, "analytic": an, "gpu_path_steps_per_s": paths * 100 / tg}
    if paths == 100_000:
        t0 = time.perf_counter(); vc, _ = cpu.bs_call(paths); row.update(cpu_s=time.perf_counter() - t0, cpu_value=vc)
    out[f"config1_black_scholes_{paths}"] = row
    print(f"config 1  BS call {paths:>8} paths x 100 steps (incl. MT19937+ICDF generation): GPU {tg * 1e3:8.2f} ms  value {vg:.6f} (analytic {an:.6f})"
          + (f"  CPU 1 thread {row['cpu_s'] * 1e3:8.1f} ms" if 'cpu_s' in row else ""))

paths = 1_000_000
mg = gpu.lmm(paths)
mg.simulate(); capi.check(L.fmc_sync())
spec = (10, 30, 2, 40, 0.02)
tb, vb = timed(lambda: mg.bermudan(*spec))
ts, _ = timed(lambda: mg.simulate())
mc = cpu.lmm(20_000); t0 = time.perf_counter(); mc.simulate(); tcs = time.perf_counter() - t0
t0 = time.perf_counter(); vcb = mc.bermudan(*spec); tcb = time.perf_counter() - t0
out["config3_bermudan_1m"] = {"gpu_simulate_s": ts, "gpu_bermudan_s": tb, "gpu_value": vb, "cpu_paths": 20_000, "cpu_simulate_s": tcs, "cpu_bermudan_s": tcb, "cpu_value": vcb}
print(f"config 3  LMM 80x80 simulate {paths} paths: GPU {ts * 1e3:8.2f} ms;  Bermudan swaption (11 exercise dates, k=6 regression): GPU {tb * 1e3:8.2f} ms  value {vb:.6f}")
print(f"          CPU 1 thread at {mc.n_paths} paths: simulate {tcs * 1e3:8.1f} ms, Bermudan {tcb * 1e3:8.1f} ms  value {vcb:.6f}  (scaled to 1M paths: {tcs * 50:.1f} s + {tcb * 50:.1f} s)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_1_3.json"), "w"), indent=1)
