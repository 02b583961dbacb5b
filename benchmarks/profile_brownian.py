#!/usr/bin/env python
"""Brownian generation only (for `ncu`): 1 Mi paths x 80 steps, three motions of the same seed (the jump states of the 2nd and 3rd come from the cache)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402

fc.ensure_init()
td = fc.TimeDiscretization(0.0, 80, 0.5)
for i in range(3):
    bm = fc.BrownianMotionCuda(td, 1, 1 << 20, 31415)
    x = bm.getBrownianIncrement(0, 0).getAverage()
    del bm
print("ok", x)
