#!/usr/bin/env python
"""BASELINE.json configs 1, 3 and 4 as wall-clock timings on one GPU next to the CPU restatement (single thread):
 1: Black-Scholes European call, Euler Monte Carlo, 100 steps, 100k and 1M paths (MonteCarloBlackScholesModelTest constants)
 3: Bermudan swaption under the LMM with conditional-expectation regression + choose(), 1M paths (CPU: bounded sample).
 4: Black-Scholes delta and vega by one reverse sweep of RandomVariableDifferentiableAAD over the GPU type, 1M paths x 100 steps."""
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
# ---- config 4: AAD delta / vega (MonteCarloBlackScholesModelTest constants), primal + adjoint sweep ----
import math  # noqa: E402

n4, steps4, T4 = 1_000_000, 100, 2.0
S0, r, sigma, K = 1.0, 0.05, 0.30, 1.05
dt4 = T4 / steps4
bm4 = fc.BrownianMotionCuda(fc.TimeDiscretization(0.0, steps4, dt4), 1, n4, 31415)
bm4.getBrownianIncrement(0, 0)
fac4 = fc.RandomVariableDifferentiableAADFactory(fc.RandomVariableCudaFactory())


def aad_run():
    s0, sig = fac4.createRandomVariable(0.0, S0), fac4.createRandomVariable(0.0, sigma)
    x = s0.log()
    drift = sig.squared().mult(-0.5).add(r).mult(dt4)
    for t in range(steps4):
        x = x.add(drift).add(sig.mult(bm4.getBrownianIncrement(t, 0)))
    V = x.exp().sub(K).floor(0.0).mult(math.exp(-r * T4)).average()
    g = V.getGradient()
    return V.doubleValue(), g[s0.getID()].getAverage(), g[sig.getID()].getAverage()


k0 = fc.stats()["n_tape_kernels"]
ta, (v4, delta4, vega4) = timed(aad_run)
kernels4 = (fc.stats()["n_tape_kernels"] - k0) // 4
d1 = (math.log(S0 / K) + (r + 0.5 * sigma * sigma) * T4) / (sigma * math.sqrt(T4))
ncdf = lambda z: 0.5 * (1.0 + math.erf(z / math.sqrt(2.0)))  # noqa: E731
out["config4_bs_aad_1m"] = {"gpu_s": ta, "value": v4, "delta": delta4, "vega": vega4, "delta_analytic": ncdf(d1),
                            "vega_analytic": S0 * math.sqrt(T4) * math.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi), "tape_kernels": kernels4}
print(f"config 4  BS delta/vega by AAD, {n4} paths x {steps4} steps: GPU {ta * 1e3:8.2f} ms ({kernels4} interpreter launches)  "
      f"delta {delta4:.5f} (analytic {ncdf(d1):.5f})  vega {vega4:.5f} (analytic {out['config4_bs_aad_1m']['vega_analytic']:.5f})")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_1_3.json"), "w"), indent=1)
