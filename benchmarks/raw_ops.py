#!/usr/bin/env python
"""Raw RandomVariable op-chain / reduction / regression / MT19937+ICDF sweep against the HBM roofline
(BASELINE.json config 5). Every case is timed with CUDA events on the runtime's
compute stream (fmc_timer_start/stop bracket exactly the flush); achieved GB/s = ALGORITHMIC bytes / time.

usage: python benchmarks/raw_ops.py [--sizes 1048576,16777216,67108864] [--cases all|b1,b2,...] [--out gpurun_out/raw_ops.json]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import finmath_cuda as fc  # noqa: E402
from finmath_cuda import _capi as capi  # noqa: E402


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, repeats=5, warmup=2):
    """fn() records the work and returns something whose evaluation is forced inside the timer by fn itself."""
    ts = []
    for i in range(warmup + repeats):
        capi.check(capi.load().fmc_sync())
        capi.timer_start()
        fn()
        capi.check(capi.load().fmc_flush())
        ms = capi.timer_stop()
        if i >= warmup:
            ts.append(ms)
    return float(np.median(ts)), float(min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1048576,16777216,67108864")
    ap.add_argument("--cases", default="all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "raw_ops.json"))
    args = ap.parse_args()
    sizes = [int(s) for s in args.sizes.split(",")]
    want = args.cases.split(",")
    fc.ensure_init()
    peak = peak_gbs()
    results = []

    def report(case, n, bytes_per_elt, ms_med, ms_min, note=""):
        gbs = bytes_per_elt * n / (ms_med * 1e-3) / 1e9
        row = {"case": case, "n": n, "bytes_per_elt": bytes_per_elt, "ms": ms_med, "ms_min": ms_min, "GBps": gbs, "frac_of_measured_peak": gbs / peak,
               "frac_of_8TBps": gbs / 8000.0, "note": note}
        results.append(row)
        print(f"{case:42s} n={n:>10d} {bytes_per_elt:3d} B/elt  {ms_med:9.4f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of measured peak {note}", flush=True)

    def sel(name):
        return "all" in want or name in want

    for n in sizes:
        rng = np.random.RandomState(31415)
        xs = [fc.RandomVariableCuda(0.0, rng.random_sample(n).astype(np.float32)) for _ in range(4)]
        x, y, z, w = xs
        l2note = "" if 4 * n * 3 > 126e6 else "(L2 resident)"

        if sel("b1"):   # unfused single ops, each materialised
            keep = []
            for name, f, b in (("add(scalar)", lambda: x.add(1.0 / 3.0), 8), ("mult(scalar)", lambda: x.mult(3.1415), 8),
                               ("div(scalar)", lambda: x.div(3.1415), 8), ("exp", lambda: x.exp(), 8), ("log", lambda: x.log(), 8),
                               ("sqrt", lambda: x.sqrt(), 8), ("add(vec)", lambda: x.add(y), 12), ("mult(vec)", lambda: x.mult(y), 12),
                               ("div(vec)", lambda: x.div(y), 12), ("accrue", lambda: x.accrue(y, 0.5), 12), ("discount", lambda: x.discount(y, 0.5), 12),
                               ("addProduct(vec,vec)", lambda: x.addProduct(y, z), 16), ("choose", lambda: x.sub(0.5).choose(y, z), 16)):
                def run(f=f):
                    keep.clear(); keep.append(f())
                ms, mn = timed(run)
                report("B1 " + name, n, b, ms, mn, l2note)
            keep.clear()

        if sel("b2"):   # fused chains
            keep = []

            def bs_step():   # X' = X + mu*dt + sigma*dW ; S = exp(X')  (2 leaves read, 2 results stored)
                keep.clear()
                xn = x.add(0.005).addProduct(y, 0.3)
                keep.extend([xn, xn.exp()])
            ms, mn = timed(bs_step)
            report("B2 BS-Euler step (X', S=exp)", n, 16, ms, mn, l2note)

            def payoff_avg():  # S.sub(K).floor(0).div(N_T).mult(N_0).getAverage(): 1 leaf read, nothing stored
                return x.sub(0.5).floor(0.0).div(1.1).mult(1.0).getAverage()
            ms, mn = timed(payoff_avg)
            report("B2 payoff chain -> getAverage (fused)", n, 4, ms, mn, l2note)

            def lmm_component():  # one LMM component update: discount, mult, add, addProduct, Euler update (App. C)
                keep.clear()
                m = fc.RandomVariableCuda(0.5).discount(x, 0.5)
                cov = z.add(m.mult(0.01))
                drift = w.addProduct(cov, 0.01)
                keep.extend([cov, x.addProduct(drift, 0.5).addProduct(y, 0.01)])
            ms, mn = timed(lmm_component)
            report("B2 LMM component update (4 in, 2 out)", n, 24, ms, mn, l2note)

            for K in (4, 16, 64):
                def chain(K=K):
                    keep.clear()
                    c = x
                    for k in range(K // 2):
                        c = c.mult(1.0001).add(y)
                    keep.append(c)
                ms, mn = timed(chain)
                report(f"B2 chain of {K} ops (2 in, 1 out)", n, 12, ms, mn, l2note)

        if sel("b3"):   # reductions on a materialised vector
            for name, f in (("getAverage", x.getAverage), ("getVariance (single pass)", x.getVariance), ("getMin", x.getMin),
                            ("getAverage(prob)", lambda: x.getAverage(y))):
                ms, mn = timed(f)
                report("B3 " + name, n, 8 if "prob" in name else 4, ms, mn, l2note)

        if sel("ref"):  # the reference's own kernels (oracle/_ref cubin), reference geometry: the bar to beat on the same box
            from oracle import ref_kernels
            if ref_kernels.available():
                import ctypes
                dv = ctypes.c_double(); capi.check(capi.load().fmc_get_option(b"device_index", ctypes.byref(dv)))
                rk = ref_kernels.ReferenceKernels(int(dv.value))
                L = capi.load()

                def ptr(rv):
                    p = ctypes.c_void_p(); capi.check(L.fmc_vec_device_ptr(rv.handle, ctypes.byref(p))); return p.value
                h = ctypes.c_uint64(); capi.check(L.fmc_vec_alloc(n, ctypes.byref(h)))
                po = ctypes.c_void_p(); capi.check(L.fmc_vec_device_ptr(h.value, ctypes.byref(po)))
                capi.check(L.fmc_sync())
                for name, kern, a, b in (("add(scalar)", "addScalar", [("p", ptr(x)), ("f", 1.0 / 3.0)], 8), ("exp", "cuExp", [("p", ptr(x))], 8),
                                         ("add(vec)", "add", [("p", ptr(x)), ("p", ptr(y))], 12), ("div(vec)", "cuDiv", [("p", ptr(x)), ("p", ptr(y))], 12),
                                         ("accrue", "accrue", [("p", ptr(x)), ("p", ptr(y)), ("f", 0.5)], 12),
                                         ("addProduct(vec,vec)", "addProduct", [("p", ptr(x)), ("p", ptr(y)), ("p", ptr(z))], 16)):
                    ms = rk.time_ms(kern, n, a + [("p", po.value)])
                    report("REF kernel " + name, n, b, ms, ms, l2note + " reference kernel, 1024 thr/block, 1 elt/thread, kernel time only")
                # the BS-Euler step as the reference executes it: 3 kernels, every intermediate through HBM
                capi.check(L.fmc_vec_release(h.value))

        if sel("b4"):   # regression normal equations
            from finmath_cuda.conditional_expectation import normal_equations
            one = fc.RandomVariableCuda(1.0)
            full = [one, x, x.squared(), y, y.squared(), x.mult(y), z, z.squared()]
            capi.check(capi.load().fmc_sync())
            for k in (3, 6, 8):
                ms, mn = timed(lambda k=k: normal_equations(full[:k], w))
                nvec = sum(1 for b in full[:k] if not b.isDeterministic()) + 1
                report(f"B4 regression normal equations k={k}", n, 4 * nvec, ms, mn, l2note)
            del full
        del xs, x, y, z, w

    if sel("b5"):   # Brownian increments: MT19937 + AS241; 4 bytes written per increment
        for (T, F) in ((100, 1), (80, 1), (40, 6)):
            for n in sizes:
                if T * F * n > (1 << 31):
                    continue
                td = fc.TimeDiscretization(0.0, T, 0.5)
                holder = []

                def gen():
                    holder.clear()
                    bm = fc.BrownianMotionCuda(td, F, n, 31415)
                    bm.getBrownianIncrement(0, 0)
                    holder.append(bm)
                ms, mn = timed(gen, repeats=3, warmup=1)
                row_n = T * F * n
                report(f"B5 Brownian T={T} F={F} paths={n}", row_n, 4, ms, mn, f"{row_n / (ms * 1e-3) / 1e9:.2f} G increments/s")
                holder.clear()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"peak_gbs": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
