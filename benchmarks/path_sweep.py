#!/usr/bin/env python
"""BASELINE.json config 2 sizes: one LMM calibration step (80 x 80 Euler simulation + 144 ATM swaptions) at 5 k ... 1 M paths on
one GPU, next to the single-threaded CPU restatement (bounded: up to 50 k paths, linear in the path count beyond).
The reference's README.md:24-28 quotes break-even at 5 000 paths, 10x at 50 000, 20x at 100 000.
usage: python benchmarks/path_sweep.py [--out gpurun_out/path_sweep.json]"""
import argparse
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

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "path_sweep.json"))
args = ap.parse_args()
fc.ensure_init()
gpu, cpu = DriverLib(), driver()
L = capi.load()
rows = []
for paths in (5000, 10000, 20000, 50000, 100000, 200000, 500000, 1000000):
    m = gpu.lmm(paths)
    for _ in range(3):
        m.step()
    capi.check(L.fmc_sync())
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        v = m.step()
    g = (time.perf_counter() - t0) / reps
    m.close()
    row = {"paths": paths, "gpu_ms": 1e3 * g, "gpu_path_steps_per_s": paths * 80 / g}
    if paths <= 50000:
        c = cpu.lmm(paths)
        t0 = time.perf_counter(); vc = c.step(); row["cpu_1thread_ms"] = 1e3 * (time.perf_counter() - t0)
        row["max_rel_diff"] = float(max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(v, vc)))
        c.close()
        per_path = row["cpu_1thread_ms"] / paths
    else:
        row["cpu_1thread_ms_extrapolated"] = per_path * paths
    cpu_ms = row.get("cpu_1thread_ms", row.get("cpu_1thread_ms_extrapolated"))
    row["speedup_vs_1thread"] = cpu_ms / row["gpu_ms"]
    rows.append(row)
    print(f"paths {paths:8d}  GPU {row['gpu_ms']:8.2f} ms/step  CPU 1 thread {cpu_ms:10.1f} ms{'' if paths <= 50000 else ' (extrapolated)'}  x{row['speedup_vs_1thread']:7.1f}"
          + (f"  max rel diff {row['max_rel_diff']:.1e}" if "max_rel_diff" in row else ""))
json.dump(rows, open(args.out, "w"), indent=1)
