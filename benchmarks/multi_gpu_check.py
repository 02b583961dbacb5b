#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank owns a contiguous path slice of the same
seeded global vectors; reductions and regression normal equations must equal the single-process numpy values of the FULL
vectors within 1e-5 relative (north-star tolerance), for every world size.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 benchmarks/multi_gpu_check.py"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import finmath_cuda as fc  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fc.ensure_init(local)
if world > 1:
    fc.distributed.init_comm_from_torch()

n = 1_000_003                                   # ragged on purpose
rng = np.random.default_rng(7)
x = rng.standard_normal(n).astype(np.float32)
y = rng.random(n).astype(np.float32)
lo, hi = fc.distributed.path_slice(n, rank, world)
X, Y = fc.RandomVariableCuda(0.0, x[lo:hi].astype(np.float64)), fc.RandomVariableCuda(0.0, y[lo:hi].astype(np.float64))
z = (x * y + np.float32(0.25)).astype(np.float32)
Z = X.mult(Y).add(0.25)
checks = {
    "average(leaf)": (X.getAverage(), x.astype(np.float64).mean()),
    "average(chain)": (Z.getAverage(), z.astype(np.float64).mean()),
    "variance(chain)": (Z.getVariance(), z.astype(np.float64).var()),
    "sampleVariance(leaf)": (X.getSampleVariance(), x.astype(np.float64).var(ddof=1)),
    "min(chain)": (Z.getMin(), float(z.min())),
    "max(leaf)": (X.getMax(), float(x.max())),
    "average(leaf, prob)": (X.getAverage(Y), (x.astype(np.float64) * y.astype(np.float64)).sum() / n),
    # order statistics of the SHARDED vector: radix-select histograms are all-reduced, the selected element is exact
    "quantile(0.95)": (X.getQuantile(0.95), float(np.sort(x)[min(max(int(math.floor((n + 1) * 0.95 - 1.0 + 0.5)), 0), n - 1)])),
    "quantile(0.0)": (X.getQuantile(0.0), float(x.min())),
    "quantileExpectation(0.1,0.9)": (X.getQuantileExpectation(0.1, 0.9), float(np.sort(x).astype(np.float64)[
        min(max(int(math.floor((n + 1) * 0.1 - 0.5)), 0), n - 1):min(max(int(math.floor((n + 1) * 0.9 - 0.5)), 0), n - 1) + 1].mean())),
    "histogram[2]": (float(X.getHistogram(np.array([-1.0, 0.0, 1.0]))[2]), float(((x > 0.0) & (x <= 1.0)).sum()) / n),
}
# conditional-expectation regression on the sharded vectors: normal equations all-reduced, coefficients solved on every rank
from finmath_cuda.conditional_expectation import MonteCarloConditionalExpectationRegression, normal_equations  # noqa: E402
one = fc.RandomVariableCuda(0.0, 1.0)
basis = [one, X, X.squared(), Y]
XtX, XtY = normal_equations(basis, Z)
xd, yd, zd = x.astype(np.float64), y.astype(np.float64), z.astype(np.float64)
B = np.stack([np.ones(n), xd, (x * x).astype(np.float64), yd])
checks["regression XtX[1][2]"] = (float(XtX[1, 2]), float((B[1] * B[2]).mean()))
checks["regression XtX[3][3]"] = (float(XtX[3, 3]), float((B[3] * B[3]).mean()))
checks["regression XtY[3]"] = (float(XtY[3]), float((B[3] * zd).mean()))
coef = MonteCarloConditionalExpectationRegression(basis).getLinearRegressionParameters(Z)
want_coef = np.linalg.lstsq((B @ B.T) / n, (B @ zd) / n, rcond=1e-10)[0]
checks["regression coefficient[3]"] = (float(coef[3]), float(want_coef[3]))
ok = True
for name, (got, want) in checks.items():
    good = abs(got - want) <= 1e-5 * max(abs(want), 1e-12)
    ok &= good
    if rank == 0:
        print(f"{name:24s} got {got:.12g} want {want:.12g} {'ok' if good else 'MISMATCH'}")
if rank == 0:
    print("multi-GPU parity", "PASSED" if ok else "FAILED", f"(world size {world})")
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
