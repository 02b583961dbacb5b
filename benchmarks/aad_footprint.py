#!/usr/bin/env python
"""Device footprint of an LMM swaption vega by RandomVariableDifferentiableAAD over RandomVariableCuda (BASELINE config 4):
the retention policy of finmath_cuda/differentiable.py ("needed": a node keeps only what its derivative rule reads) against the
store-everything tree ("all"), with and without releasing retained values during the reverse sweep.
usage: python benchmarks/aad_footprint.py [paths] [periods]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
import finmath_cuda.differentiable as D  # noqa: E402

fc.ensure_init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 40
delta, L0, sigma0, strike = 0.5, 0.02, 0.006, 0.02
td = fc.TimeDiscretization(0.0, NP, delta)
bm = fc.BrownianMotionCuda(td, 1, n, 31415)
plain = fc.RandomVariableCudaFactory()
for t in range(NP):
    bm.getBrownianIncrement(t, 0)


def vega(retention, release):
    D.RETENTION = retention
    fc.sync(); fc.pool_trim(); fc.reset_stats()
    base = fc.stats()["bytes_in_use"]
    t0 = time.perf_counter()
    sig = fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, sigma0))
    libor = [plain.createRandomVariable(0.0, L0) for _ in range(NP)]
    exercise, at_ex = NP // 2, None
    for t in range(NP):
        if t == exercise:
            at_ex = list(libor)
        dW, acc, new = bm.getBrownianIncrement(t, 0), None, list(libor)
        for i in range(t + 1, NP):
            tr = sig.mult(libor[i].mult(delta).add(1.0).invert().mult(delta))
            acc = tr if acc is None else acc.add(tr)
            new[i] = libor[i].add(acc.mult(sig).mult(delta)).add(sig.mult(dW))
        libor = new
    value = None
    for i in range(NP - 1, exercise - 1, -1):
        payoff = at_ex[i].sub(strike).mult(delta)
        value = payoff if value is None else value.add(payoff)
        value = value.discount(at_ex[i], delta)
    V = value.floor(0.0).average()
    del libor, new, at_ex, value, acc, tr, payoff
    g = V.getGradient(release=release)[sig.getID()].getAverage()
    fc.sync()
    return {"retention": retention, "release_during_sweep": release, "vega": g, "value": V.doubleValue(), "seconds": time.perf_counter() - t0,
            "peak_device_bytes": fc.stats()["bytes_high_water"] - base, "kernels": fc.stats()["n_tape_kernels"]}


rows = [vega("all", False), vega("needed", False), vega("needed", True)]
D.RETENTION = "needed"
for r in rows:
    print(json.dumps(r))
print(json.dumps({"paths": n, "periods": NP, "footprint_ratio_needed_vs_all": rows[1]["peak_device_bytes"] / max(1, rows[0]["peak_device_bytes"]),
                  "footprint_ratio_release_vs_all": rows[2]["peak_device_bytes"] / max(1, rows[0]["peak_device_bytes"])}))
