#!/usr/bin/env python
"""Cold and warm cost of BrownianMotionCuda generation (1 Mi paths x 80 steps): first motion of the process, a new seed with the
same layout (jump-ahead of the block start states), the same seed again (cached states)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402

fc.ensure_init()
td = fc.TimeDiscretization(0.0, 80, 0.5)
for label, seed in (("first motion of the process", 31415), ("new seed, same layout", 31416), ("same seed again", 31416), ("third seed", 7), ("third seed again", 7)):
    fc.sync(); t0 = time.perf_counter()
    bm = fc.BrownianMotionCuda(td, 1, 1 << 20, seed)
    inc = bm.getBrownianIncrement(0, 0)
    fc.sync(); dt = time.perf_counter() - t0
    print(f"{label:32s} {1e3 * dt:8.2f} ms")
    del bm, inc
