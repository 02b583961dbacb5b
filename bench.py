#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: "LMM ATM calibration sec & path-steps/s at 1/2/4/8
B200; op HBM GB/s vs peak").

A "step" is one pass of the calibration inner loop of LIBORMarketModelCalibrationATMTest (T-ATM): Euler simulation
of the LIBOR market model (80 time steps x 80 forward rates, 1 factor, seed 31415, SPOT measure, NORMAL state
space; every arithmetic step one RandomVariable call) followed by the valuation of the 144 in-horizon ATM calibration
swaptions (each ending in getAverage()), all through RandomVariableCudaFactory / BrownianMotionCuda over the C ABI.

  value : path-steps/s (paths x 80 time steps / step time), Brownian increments already resident in HBM
  e2e   : same, but the increments come from HOST double arrays through createRandomVariable(time, double[]) inside
          the timed region (what T-ATM:283 does with BrownianMotionFromMersenneRandomNumbers + RandomVariableCudaFactory)
          and the 144 results are read back to the host
  roofline      : the op-tape interpreter kernel, algorithmic bytes / CUDA-event time, vs measured HBM peak
  cpu_baseline  : the CPU oracle (C++ restatement of RandomVariableFromFloatArray + the same driver source) on a
                  bounded sample, 1 thread
  --impl reference : the reference's CPU path (same oracle build; no JVM exists here) on all host threads

N > 1: launched by torchrun, one process per GPU; every rank owns a contiguous path slice (weak scaling: paths per
GPU fixed), the only exchange is the all-reduce of reduction partials inside the runtime (NCCL).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

N_PERIODS = 80          # T-ATM:275-278: 0..40y in 0.5y steps
DELTA = 0.5
SEED = 31415            # T-ATM:283
METRIC = "lmm_atm_path_steps_per_s"
UNIT = "path-steps/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t0 = self.t1 = None

    def start(self):
        """Started BEFORE the warm-up (nvidia-smi needs a few hundred ms to come up); begin()/end() bracket the timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def begin(self): self.t0 = time.perf_counter()
    def end(self): self.t1 = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)                      # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        note = None
        if not inside and self.lines:         # region shorter than the sampling interval: the samples closest to it
            mid = 0.5 * ((self.t0 or 0.0) + (self.t1 or 0.0))
            inside = [ln for (_, ln) in sorted(self.lines, key=lambda tl: abs(tl[0] - mid))[:2]]
            note = "timed region shorter than the 50 ms sampling interval: nearest samples"
        sm, smax, reasons = [], None, set()
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def run_reference(args) -> None:
    """The reference's CPU implementation of the path (RandomVariableFromFloatArray restated in C++; no JVM here),
    every host thread simulating its own path slice; same config / metric / unit as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle.workloads_oracle import driver
    lib = driver()
    cores = os.cpu_count() or 1
    paths_per_thread = args.ref_paths_per_thread
    total = cores * paths_per_thread
    models = [lib.lmm(total, N_PERIODS, DELTA, 1, SEED, 0, (i * paths_per_thread, (i + 1) * paths_per_thread)) for i in range(cores)]
    for m in models:
        m.use_market_curve()
    pool = ThreadPoolExecutor(cores)

    def step():
        # each worker values its slice; the slice averages are combined with the slice weights (equal sizes)
        vals = list(pool.map(lambda m: m.step(), models))     # ctypes releases the GIL: the C++ drivers run in parallel
        return sum(vals) / len(vals)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = total * N_PERIODS * args.steps / dt
    # the same sample on finmath-lib's default CPU type, RandomVariableFromDoubleArray (north star: "DoubleArray/FloatArray path")
    from oracle.workloads_oracle import driver_f64
    lib64 = driver_f64()
    models64 = [lib64.lmm(total, N_PERIODS, DELTA, 1, SEED, 0, (i * paths_per_thread, (i + 1) * paths_per_thread)) for i in range(cores)]
    for m in models64:
        m.use_market_curve()
    list(pool.map(lambda m: m.step(), models64))
    t0 = time.perf_counter()
    list(pool.map(lambda m: m.step(), models64))
    value64 = total * N_PERIODS / (time.perf_counter() - t0)
    # the calibration as the reference test runs it: 10 000 paths, ONE thread (T-ATM:154,319), for a fixed budget of 2 iterations
    calibration = None
    if not args.no_calibration:
        cm = lib.lmm(10000, N_PERIODS, DELTA, 1, SEED, 0)
        cm.use_market_curve()
        r2 = cm.calibrate(max_iterations=2)
        calibration = {"t_atm_10k_paths_2_iterations": {"evaluations": r2["evaluations"], "seconds": r2["seconds"], "seconds_per_evaluation": r2["seconds_per_evaluation"],
                                                        "cores": 1, "note": "single thread, like numberOfThreads = 1 in T-ATM:319"}}
        cm.close()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "LIBORMarketModelCalibrationATMTest inner loop: LMM Euler simulation 80x80, 1 factor + 144 ATM swaptions",
                   "paths": total, "time_steps": N_PERIODS, "note": "bounded sample of the workload on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{total} paths ({paths_per_thread} per thread x {cores} threads), {args.steps} steps; C++ restatement of "
                                   "RandomVariableFromFloatArray (no JVM in this environment)"},
        "cpu_double_array": {"value": value64, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{total} paths, 1 step; C++ restatement of finmath-lib's RandomVariableFromDoubleArray (not part of the ratio)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "calibration": calibration,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (an LMM Euler-step launch of the tape
# interpreter at 1 Mi paths) from one `ncu --set full` capture; NOT measured in this run, NOT comparable with the per-step
# algorithmic bytes: reported under its own name with the capture it came from (None until this round's capture exists).
NCU_DRAM_BYTES_DOMINANT_LAUNCH = 869_150_976     # 245.40 MB read + 623.75 MB written, first launch of profiles/prof_sim_r5.txt (226.4 us): one
                                                 # window of three Euler time steps (round 1 / early round 2: 533 MB for ONE time step)
NCU_ALGORITHMIC_BYTES_OF_THAT_LAUNCH = 1_266_679_808   # its tape: 78 leaf vectors (75 rates + 3 Brownian increments) + 224 result vectors x 4 MiB.
                                                 # The DRAM traffic is BELOW it: part of the stores still sits in the L2 when the kernel ends and
                                                 # some leaves are hit there; the rates a step stores are no longer read back inside the window
NCU_CAPTURE_FILE = "profiles/prof_sim_r5.txt (ncu --set full --clock-control none -k regex:tape_kernel -s 173 -c 3 of bench.py --steps 2 --warmup 1; launch list profiles/launches_r5.txt)"

PARITY_REL_TOL = 1e-4       # north star: "Monte-Carlo prices ... match within 1e-4 relative on identical seeds"


def rel_diff(got, want):
    import numpy as np
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12)))


def _chain16(x, y):
    c = x
    for k in range(4):
        c = c.mult(1.0001).add(y).sub(0.001).mult(y)
    return c


def run_ours(args) -> None:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import finmath_cuda as fc
    from finmath_cuda import _capi as capi
    from finmath_cuda.workloads import DriverLib

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fc.ensure_init(local_rank)
    if world > 1:
        fc.distributed.init_comm_from_torch()

    paths_per_gpu = args.paths
    total_paths = paths_per_gpu * world                      # weak scaling: per-GPU work fixed
    p0, p1 = rank * paths_per_gpu, (rank + 1) * paths_per_gpu
    lib = DriverLib()
    model = lib.lmm(total_paths, N_PERIODS, DELTA, 1, SEED, 0, (p0, p1))
    model.use_market_curve()                                  # T-ATM:526-663 swap curve (idealised schedules), not a synthetic one
    if args.valuation_threads != 1:
        if world > 1:
            raise SystemExit("--valuation-threads > 1 is a single-rank option: sharded ranks must issue their reductions in one order")
        model.set_valuation_threads(args.valuation_threads)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ----
    values = None
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        values = model.step()
    capi.check(capi.load().fmc_reset_stats())
    capi.set_option("profile", 1)
    capi.profile_read()
    barrier()
    sampler.begin()
    capi.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        values = model.step()
    dev_ms = capi.timer_stop()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sampler.end()
    clocks = sampler.stop()
    touched = ctypes.c_double()
    capi.check(capi.load().fmc_get_option(b"profile_touched_bytes", ctypes.byref(touched)))
    prof = capi.profile_read()
    capi.set_option("profile", 0)
    st = fc.stats()
    host_prof = {}
    for key in ("host_us_codegen", "host_us_launch", "host_us_sync"):
        v = ctypes.c_double()
        capi.check(capi.load().fmc_get_option(key.encode(), ctypes.byref(v)))
        host_prof[key.replace("host_us_", "") + "_ms_per_step"] = v.value / 1e3 / args.steps
    step_ms = max_over_ranks(max(dev_ms, 0.0)) / args.steps
    value = total_paths * N_PERIODS / (step_ms * 1e-3)

    # ---- end-to-end arm: Brownian increments enter from host double arrays inside the timed region ----
    host_bytes = model.prepare_host_brownian()
    for _ in range(max(1, min(args.warmup, 2))):
        model.step(from_host=True)
    capi.check(capi.load().fmc_reset_stats())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        values_e2e = model.step(from_host=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    st2 = fc.stats()
    e2e_ms = max_over_ranks(1e3 * e2e_s) / args.steps
    e2e_value = total_paths * N_PERIODS / (e2e_ms * 1e-3)

    # ---- the same with the caller's doubles in PINNED host memory (fmc_host_alloc) and the asynchronous upload
    # fmc_vec_from_f64_pinned: DMA on the copy stream + (float) cast on the device, overlapping the kernels. An extension for
    # callers that can pin their arrays; the `e2e` key above stays the unmodified createRandomVariable(time, double[]) path.
    for _ in range(max(1, min(args.warmup, 2))):
        model.step(from_host=2)
    capi.check(capi.load().fmc_reset_stats())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        values_pinned = model.step(from_host=2)
    torch.cuda.synchronize()
    pinned_s = time.perf_counter() - t0
    barrier()
    st3 = fc.stats()
    pinned_ms = max_over_ranks(1e3 * pinned_s) / args.steps

    # ---- the same step for PRICE products (ValueUnit.VALUE): the objective function collects all 144 value vectors and averages
    # them in a second loop; the runtime sums the vectors one flush produced in one launch. Single rank only; not the headline:
    # T-ATM calibrates to implied volatilities, whose products take their own average one by one (T-ATM:261,511).
    price_products = None
    if world == 1:
        model.set_price_products(True)
        for _ in range(max(1, min(args.warmup, 3))):
            values_pp = model.step()
        capi.timer_start()
        for _ in range(args.steps):
            values_pp = model.step()
        pp_ms = capi.timer_stop() / args.steps
        model.set_price_products(False)
        rel = rel_diff(values_pp, values)
        price_products = {"ms_per_step": pp_ms, "value": total_paths * N_PERIODS / (pp_ms * 1e-3), "unit": UNIT,
                          "max_rel_diff_vs_headline_values": rel, "ok": bool(rel <= 1e-9),
                          "what": "all product value vectors first, their averages in a second loop (AbstractLIBORCovarianceModelParametric's "
                                  "objective function with ValueUnit.VALUE products); batched averages in the runtime"}
        if not price_products["ok"]:
            raise SystemExit(f"bench: price-product step differs from the headline step: {price_products}")

    # ---- the calibration itself (BASELINE.json metric: "LMM ATM calibration sec"): Levenberg-Marquardt with the settings of
    # T-ATM:317-340 on the bench's model for a fixed budget of iterations, and on T-ATM's own 10 000 paths to convergence ----
    calibration = None
    if not args.no_calibration:
        p_init = model.parameters()
        barrier()
        r = model.calibrate(max_iterations=args.calibration_iterations)
        calibration = {"paths_per_gpu": paths_per_gpu, "paths_total": total_paths, "lm_iterations": r["iterations"], "evaluations": r["evaluations"],
                       "seconds": max_over_ranks(r["seconds"]), "seconds_per_evaluation": r["seconds_per_evaluation"],
                       "rms_error": r["rms_error"], "mean_deviation": r["mean_deviation"],
                       "settings": "Levenberg-Marquardt, lambda 0.1, accuracy 1e-7, parameter step 1e-4, 48 parameters, 144 products, 1 thread (T-ATM:317-340)"}
        model.set_parameters(p_init)
        if world == 1:
            cm = lib.lmm(10000, N_PERIODS, DELTA, 1, SEED, 0, (0, 10000))
            cm.use_market_curve()
            r2 = cm.calibrate(max_iterations=2)
            calibration["t_atm_10k_paths_2_iterations"] = {"evaluations": r2["evaluations"], "seconds": r2["seconds"], "seconds_per_evaluation": r2["seconds_per_evaluation"]}
            r3 = cm.calibrate(max_iterations=200)
            calibration["t_atm_10k_paths_to_convergence"] = {"iterations": r2["iterations"] + r3["iterations"], "evaluations": r2["evaluations"] + r3["evaluations"],
                                                             "seconds": r2["seconds"] + r3["seconds"], "mean_deviation": r3["mean_deviation"],
                                                             "rms_error": r3["rms_error"], "reference_bound": "|mean deviation| < 2e-4 (T-ATM:466)",
                                                             "ok": bool(abs(r3["mean_deviation"]) < 2e-4)}
            cm.close()

    # ---- the other configurations of BASELINE.json, each as a compact extra key of the same line ----
    extras = {}
    if not args.no_extras:
        import numpy as np
        # config 3: Bermudan swaption (11 exercise dates, k = 6 regression, choose) at 1 Mi paths IN TOTAL, sharded over the ranks (strong scaling)
        bp = 1 << 20
        per = ((bp + world - 1) // world + 3) // 4 * 4
        b0, b1 = min(bp, rank * per), min(bp, (rank + 1) * per)
        bm = lib.lmm(bp, N_PERIODS, DELTA, 1, SEED, 0, (b0, b1))
        bm.use_market_curve()
        spec = (10, 30, 2, 40, 0.02)
        bm.simulate(); bv = bm.bermudan(*spec); barrier()
        t_sim, t_val = [], []
        for _ in range(3):
            barrier(); t0 = time.perf_counter(); bm.simulate(); capi.check(capi.load().fmc_sync()); t1 = time.perf_counter()
            bv = bm.bermudan(*spec); t2 = time.perf_counter()
            t_sim.append(max_over_ranks(t1 - t0)); t_val.append(max_over_ranks(t2 - t1))
        extras["bermudan_1m_paths_sharded"] = {"paths_total": bp, "ranks": world, "simulate_ms": 1e3 * min(t_sim), "valuation_ms": 1e3 * min(t_val), "value": bv,
                                               "what": "BASELINE config 3: backward induction, regression on 6 basis functions at 11 exercise dates, choose(); strong scaling"}
        bm.close()
        if world == 1:
            # config 2 path sweep (5 k ... 500 k paths; the 1 Mi point is the headline above): one calibration step, device-resident
            sweep = {}
            for pth in (5000, 10000, 20000, 50000, 100000, 200000, 500000):
                sm_ = lib.lmm(pth, N_PERIODS, DELTA, 1, SEED, 0, (0, pth))
                sm_.use_market_curve()
                for _ in range(3):
                    sm_.step()
                capi.timer_start()
                for _ in range(5):
                    sm_.step()
                sweep[str(pth)] = capi.timer_stop() / 5
                sm_.close()
            extras["path_sweep_ms_per_step"] = sweep
            # config 5: raw operations at 1e8 elements against the HBM roofline (algorithmic bytes / CUDA-event time of the flush)
            n_raw = 100_000_000
            rng = np.random.RandomState(SEED)
            xs = [fc.RandomVariableCuda(0.0, rng.random_sample(n_raw).astype(np.float32)) for _ in range(3)]
            x, y, z = xs
            peak_ = float(measured_peaks()[0].get("hbm_gbs", 6650.0))

            def raw(fn, bytes_per_elt):
                ts = []
                for i in range(5):
                    capi.check(capi.load().fmc_sync()); capi.timer_start(); r = fn(); capi.check(capi.load().fmc_flush()); ts.append(capi.timer_stop()); del r
                ms = float(np.median(ts[2:]))
                g = bytes_per_elt * n_raw / (ms * 1e-3) / 1e9
                return {"ms": ms, "GBps": g, "frac_of_measured_peak": g / peak_}
            ops = {"add(vec)": (lambda: x.add(y), 12), "mult(scalar)": (lambda: x.mult(3.1415), 8), "accrue": (lambda: x.accrue(y, 0.5), 12),
                   "discount": (lambda: x.discount(y, 0.5), 12), "addProduct(vec,vec)": (lambda: x.addProduct(y, z), 16), "exp": (lambda: x.exp(), 8),
                   "log": (lambda: x.log(), 8), "chain of 16 ops": (lambda: _chain16(x, y), 12),
                   "getAverage": (lambda: x.getAverage(), 4), "getVariance": (lambda: x.getVariance(), 4),
                   "payoff chain -> getAverage": (lambda: x.sub(0.5).floor(0.0).div(1.1).mult(0.9).getAverage(), 4)}
            extras["raw_ops_1e8"] = {k: raw(f, b) for k, (f, b) in ops.items()}
            del xs, x, y, z
            fc.pool_trim(); capi.check(capi.load().fmc_sync())   # the 400 MB blocks go back to the driver outside the timings below
            # Brownian generation (MT19937 + inverse normal), 1 Mi paths x 80 steps
            td = fc.TimeDiscretization(0.0, N_PERIODS, DELTA)
            tsb = []
            for i in range(4):
                capi.check(capi.load().fmc_sync()); t0 = time.perf_counter()
                bmo = fc.BrownianMotionCuda(td, 1, 1 << 20, SEED + 1)
                inc = bmo.getBrownianIncrement(0, 0); capi.check(capi.load().fmc_sync()); tsb.append(time.perf_counter() - t0); del bmo, inc
            extras["brownian_1m_x_80"] = {"first_ms": 1e3 * tsb[0], "cached_jump_ms": 1e3 * min(tsb[1:]), "increments_per_s": (1 << 20) * 80 / min(tsb[1:]),
                                          "GBps_written": 4 * (1 << 20) * 80 / min(tsb[1:]) / 1e9}

    # ---- the workload at the parity sample's size, on every rank's slice of the SAME global paths (compared with the oracle below) ----
    sample_values = None
    if not args.no_cpu_baseline:
        sp = args.cpu_sample_paths if world == 1 else args.multi_gpu_check_paths
        per = (sp + world - 1) // world
        per = (per + 3) // 4 * 4                                  # slice boundaries on multiples of 4 elements
        s0, s1 = min(sp, rank * per), min(sp, (rank + 1) * per)
        sm = lib.lmm(sp, N_PERIODS, DELTA, 1, SEED, 0, (s0, s1))
        sm.use_market_curve()
        sample_values = sm.step()
        sm.close()
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    import numpy as np
    peaks, peak_src = measured_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = prof["tape_algorithmic_bytes"] / (prof["tape_ms"] * 1e-3) / 1e9 if prof["tape_ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "tape_kernel (op-tape interpreter)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": None,
                "launches": prof["tape_launches"], "kernel_ms_per_step": prof["tape_ms"] / args.steps,
                "kernel_share_of_step": (prof["tape_ms"] / args.steps) / step_ms if step_ms > 0 else None,
                "algorithmic_bytes_per_step": prof["tape_algorithmic_bytes"] / args.steps,
                # every vector the kernels read or wrote, INCLUDING the re-read of results an earlier kernel of the same
                # flush stored (the first time step of a window re-reads the rates the window before it wrote): what HBM actually moves
                "touched_bytes_per_step": touched.value / args.steps,
                "achieved_touched": touched.value / (prof["tape_ms"] * 1e-3) / 1e9 if prof["tape_ms"] > 0 else 0.0,
                "frac_touched": (touched.value / (prof["tape_ms"] * 1e-3) / 1e9) / peak if prof["tape_ms"] > 0 else 0.0}
    roofline["ncu_dram_bytes_dominant_launch"] = NCU_DRAM_BYTES_DOMINANT_LAUNCH
    roofline["ncu_algorithmic_bytes_of_that_launch"] = NCU_ALGORITHMIC_BYTES_OF_THAT_LAUNCH
    roofline["ncu_capture"] = NCU_CAPTURE_FILE

    # ---- parity of what was timed: the CPU oracle on a bounded sample, the GPU on exactly the same paths ----
    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle.workloads_oracle import driver
        olib = driver()
        sample_paths = args.cpu_sample_paths if world == 1 else args.multi_gpu_check_paths
        om = olib.lmm(sample_paths, N_PERIODS, DELTA, 1, SEED, 0)
        om.use_market_curve()
        t0 = time.perf_counter()
        ovalues = om.step()
        dt = time.perf_counter() - t0
        del om
        d = rel_diff(sample_values, ovalues)
        parity = {"paths": sample_paths, "products": int(len(ovalues)), "max_rel_diff": d, "tol": PARITY_REL_TOL, "ok": bool(d <= PARITY_REL_TOL),
                  "what": "all swaption values of one step, CUDA path vs CPU oracle (RandomVariableFromFloatArray restated) on the same paths and seed"}
        if world == 1:
            cpu_baseline = {"value": sample_paths * N_PERIODS / dt, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"1 step at {sample_paths} paths, single thread ({dt:.1f} s); C++ restatement of RandomVariableFromFloatArray",
                            "max_abs_price_diff_vs_gpu_sample": float(np.max(np.abs(np.asarray(sample_values) - np.asarray(ovalues)))),
                            "max_rel_price_diff_vs_gpu_sample": d}
        else:
            parity["what"] += f"; the GPU run is sharded over {world} ranks (global Brownian stream by jump-ahead, reductions exchanged between the GPUs)"
    else:
        parity = None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LIBORMarketModelCalibrationATMTest inner loop: LMM Euler simulation 80x80, 1 factor + 144 ATM swaptions",
                   "paths_per_gpu": paths_per_gpu, "paths_total": total_paths, "time_steps": N_PERIODS, "seed": SEED,
                   "parallelism": f"path-sharded x{world}", "valuation_threads": args.valuation_threads, "l2": "inputs larger than L2 (simulation state >> 126 MB)",
                   "forward_curve": "EUR swap curve of T-ATM:526-663 (idealised schedules)", "wall_ms_per_step": wall_ms / args.steps},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": st2["h2d_bytes"] // args.steps,
                "d2h_bytes_per_step": st2["d2h_bytes"] // args.steps, "host_input_bytes": host_bytes,
                "input": "pageable host double[] through createRandomVariable(time, double[]) (cast on host threads into pinned staging)"},
        "e2e_pinned": {"value": total_paths * N_PERIODS / (pinned_ms * 1e-3), "unit": UNIT, "ms_per_step": pinned_ms,
                       "h2d_bytes_per_step": st3["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st3["d2h_bytes"] // args.steps,
                       "input": "pinned host double[] (fmc_host_alloc) through fmc_vec_from_f64_pinned: asynchronous DMA, cast on the device"},
        "gpu_launches": st["n_kernels"],
        "gpu_launches_per_step": st["n_kernels"] / args.steps,
        "ops_recorded_per_step": st["n_ops_recorded"] / args.steps,
        "nodes_stored_per_step": st["n_nodes_stored"] / args.steps, "nodes_fused_per_step": st["n_nodes_fused"] / args.steps,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "host_profile": host_prof,
        ("parity" if world == 1 else "multi_gpu_parity"): parity,
        "calibration": calibration, "price_products_step": price_products, "extras": extras,
        "price_check": {"first_values": [float(v) for v in values[:3]], "e2e_equal": bool((values == values_e2e).all()),
                        "e2e_pinned_equal": bool((values == values_pinned).all())},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"bench.py: parity check failed: max relative difference {parity['max_rel_diff']:.3g} > {PARITY_REL_TOL}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=1 << 20, help="paths per GPU (north-star: 1M-path runs)")
    ap.add_argument("--cpu-sample-paths", type=int, default=393216)
    ap.add_argument("--multi-gpu-check-paths", type=int, default=98304,
                    help="N > 1: total paths of the sharded run that rank 0 compares with the CPU oracle (untimed)")
    ap.add_argument("--ref-paths-per-thread", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-calibration", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Bermudan / path-sweep / raw-op extra keys")
    ap.add_argument("--calibration-iterations", type=int, default=1, help="Levenberg-Marquardt iterations of the calibration key (1 iteration = 50 simulations)")
    ap.add_argument("--valuation-threads", type=int, default=1,
                    help="host threads valuing the calibration products (the reference test uses 1, LIBORMarketModelCalibrationATMTest.java:319)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
