#!/usr/bin/env python
"""Shapes of the interpreter launches of one LMM step (FMC_LOG_TAPES): usage: python benchmarks/log_tapes.py [paths]"""
import os
import sys

os.environ["FMC_LOG_TAPES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
m.step()
