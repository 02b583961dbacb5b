#!/usr/bin/env python
"""BASELINE.json config 3: Bermudan swaption under the LMM with conditional-expectation regression, 1 M paths sharded across
the ranks (one process per GPU; run under torchrun, or alone for the single-GPU numbers). Every rank simulates its contiguous
path slice of the SAME global Brownian motion (MT19937 jump-ahead), the regression's normal equations and the final average
are the only things that cross NVLink, so the value must not depend on the number of ranks.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 benchmarks/bermudan_sharded.py
  python benchmarks/bermudan_sharded.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fc.ensure_init(local)
if world > 1:
    fc.distributed.init_comm_from_torch()

PATHS = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
SPEC = (10, 30, 2, 40, 0.02)          # exercise every 2nd period from 10 to 30, swap ends at period 40, strike 2 %
SINGLE_GPU_VALUE_1M = 0.07606907314274954   # benchmarks/configs.py on one GPU (gpurun_out/configs_1_3.json)
lo, hi = fc.distributed.path_slice(PATHS, rank, world)
m = DriverLib().lmm(PATHS, 80, 0.5, 1, 31415, 0, (lo, hi))
L = capi.load()


def barrier():
    capi.check(L.fmc_sync())
    if world > 1:
        dist.barrier()


def timed(fn, reps=3):
    fn(); barrier()
    best = 1e30; out = None
    for _ in range(reps):
        barrier(); t0 = time.perf_counter(); out = fn(); capi.check(L.fmc_sync())
        t = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best, out


ts, _ = timed(m.simulate)
tb, v = timed(lambda: m.bermudan(*SPEC))
vals = [v]
if world > 1:
    g = [None] * world
    dist.all_gather_object(g, v)
    vals = g
ok = all(abs(x - vals[0]) <= 1e-12 * abs(vals[0]) for x in vals)
if PATHS == 1_000_000:
    ok = ok and abs(v - SINGLE_GPU_VALUE_1M) <= 1e-4 * SINGLE_GPU_VALUE_1M
if rank == 0:
    row = {"ranks": world, "paths_total": PATHS, "paths_per_rank": hi - lo, "simulate_ms": 1e3 * ts, "bermudan_ms": 1e3 * tb, "value": v,
           "same_on_all_ranks": ok, "single_gpu_value_1m": SINGLE_GPU_VALUE_1M}
    print(json.dumps(row))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(row, open(os.path.join(ROOT, "gpurun_out", f"bermudan_{world}gpu.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
