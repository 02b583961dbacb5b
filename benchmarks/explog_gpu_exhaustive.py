#!/usr/bin/env python
"""exp and log of ALL 2^32 float bit patterns on the GPU (through the C ABI) against the oracle (glibc exp / log in double,
rounded to float = RandomVariableFromFloatArray.java:903-921). Prints the number of differing results per function.
usage: python benchmarks/explog_gpu_exhaustive.py [log2 of the stride, default 0 = exhaustive]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from oracle import oracle as O  # noqa: E402

fc.ensure_init()
shift = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CH = 1 << 26
t0 = time.perf_counter()
for name, op in (("exp", O.EXP), ("log", O.LOG)):
    bad = 0
    first = []
    for c in range(0, 1 << 32, CH << shift):
        bits = (np.arange(c, c + (CH << shift), 1 << shift, dtype=np.uint64)).astype(np.uint32)
        x = bits.view(np.float32)
        X = fc.RandomVariableCuda(0.0, x)
        got = getattr(X, name)().getRealizationsFloat()
        want = O.op_v(op, x)
        same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
        idx = np.flatnonzero(~same)
        bad += idx.size
        for i in idx[:3]:
            if len(first) < 8:
                first.append((hex(int(bits[i])), float(got[i]), float(want[i])))
        del X
    print(f"{name}: {(1 << 32) >> shift} inputs, {bad} results differ from the oracle {first}", flush=True)
print(f"{time.perf_counter() - t0:.0f} s")
