"""ctypes binding of the workload drivers (drivers/driver_api.cpp): Black-Scholes Euler MC, LIBOR-market-model
simulation + ATM swaption valuation (the inner loop of LIBORMarketModelCalibrationATMTest), Bermudan swaption.

`DriverLib()` binds lib/libfmdrivers_cuda.so (the product backend). The same C entry points exist in
oracle/libfmdrivers_oracle.so (CPU oracle backend, tests and CPU baseline only) — see oracle/workloads_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi as capi

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_DRIVER_LIB = os.path.normpath(os.path.join(_HERE, "..", "lib", "libfmdrivers_cuda.so"))


class DriverLib:
    def __init__(self, path: str = CUDA_DRIVER_LIB):
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: run __graft_entry__.build()")
        if path == CUDA_DRIVER_LIB:
            capi.load()            # libfmcuda.so first (rpath covers it too)
        L = C.CDLL(path)
        i64, dbl, i32, vp = C.c_int64, C.c_double, C.c_int, C.c_void_p
        dp = C.POINTER(C.c_double)
        L.fmd_last_error.restype = C.c_char_p
        L.fmd_backend.restype = i32
        L.fmd_bs_call.argtypes = [i64, i32, dbl, i32, i32, dbl, dbl, dbl, dbl, dbl, dp, dp]
        L.fmd_lmm_create.argtypes = [i64, i32, dbl, i32, i32, i32, i64, i64]
        L.fmd_lmm_create.restype = vp
        L.fmd_lmm_destroy.argtypes = [vp]
        L.fmd_lmm_num_products.argtypes = [vp]
        L.fmd_lmm_num_parameters.argtypes = [vp]
        L.fmd_lmm_get_parameters.argtypes = [vp, vp]
        L.fmd_lmm_product_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), dp, dp]
        L.fmd_lmm_prepare_host_brownian.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.fmd_lmm_step.argtypes = [vp, vp, i32, vp]
        L.fmd_lmm_get_libor.argtypes = [vp, i32, i32, vp, i64]
        L.fmd_lmm_implied_vols.argtypes = [vp, vp, vp]
        L.fmd_lmm_bermudan.argtypes = [vp, i32, i32, i32, i32, dbl, dp]
        L.fmd_lmm_simulate.argtypes = [vp]
        L.fmd_lmm_use_market_curve.argtypes = [vp]
        L.fmd_lmm_get_forward_rates.argtypes = [vp, vp]
        L.fmd_lmm_set_parameters.argtypes = [vp, vp]
        L.fmd_lmm_calibrate.argtypes = [vp, i32, dbl, dbl, dbl, vp, vp]
        L.fmd_lmm_set_valuation_threads.argtypes = [vp, i32]
        L.fmd_lmm_set_price_products.argtypes = [vp, i32]
        self.L = L

    def check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError((self.L.fmd_last_error() or b"").decode("utf-8", "replace"))

    @property
    def is_cuda(self) -> bool:
        return self.L.fmd_backend() == 1

    def bs_call(self, n_paths: int, n_steps: int = 100, dt: float = 1.0, seed: int = 31415, seed_mode: int = 0, S0: float = 1.0,
                r: float = 0.05, sigma: float = 0.30, maturity: float = 2.0, strike: float = 1.05):
        """MonteCarloBlackScholesModelTest.java:62-76 constants by default. Returns (monte-carlo value, analytic value)."""
        v, a = C.c_double(), C.c_double()
        self.check(self.L.fmd_bs_call(n_paths, n_steps, dt, seed, seed_mode, S0, r, sigma, maturity, strike, C.byref(v), C.byref(a)))
        return v.value, a.value

    def lmm(self, n_paths: int, n_periods: int = 80, delta: float = 0.5, n_factors: int = 1, seed: int = 31415, seed_mode: int = 0,
            path_range: tuple[int, int] | None = None) -> "Lmm":
        return Lmm(self, n_paths, n_periods, delta, n_factors, seed, seed_mode, path_range)


class Lmm:
    """LIBORMarketModelCalibrationATMTest.java:275-314 model: 0..40y in 0.5y steps, 80 rates, 1 factor, seed 31415."""

    def __init__(self, lib: DriverLib, n_paths, n_periods, delta, n_factors, seed, seed_mode, path_range):
        self.lib = lib
        p0, p1 = (0, -1) if path_range is None else path_range
        self.h = lib.L.fmd_lmm_create(n_paths, n_periods, delta, n_factors, seed, seed_mode, p0, p1)
        if not self.h:
            raise RuntimeError((lib.L.fmd_last_error() or b"").decode())
        self.n_paths, self.n_periods = n_paths, n_periods
        self.local_paths = n_paths if path_range is None else path_range[1] - path_range[0]
        self.n_products = lib.L.fmd_lmm_num_products(self.h)
        self.n_parameters = lib.L.fmd_lmm_num_parameters(self.h)

    def close(self) -> None:
        if self.h:
            self.lib.L.fmd_lmm_destroy(self.h)
            self.h = None

    __del__ = close

    def parameters(self) -> np.ndarray:
        out = np.empty(self.n_parameters)
        self.lib.L.fmd_lmm_get_parameters(self.h, out.ctypes.data)
        return out

    def products(self):
        res = []
        for k in range(self.n_products):
            e, m, s, v = C.c_int(), C.c_int(), C.c_double(), C.c_double()
            self.lib.L.fmd_lmm_product_info(self.h, k, C.byref(e), C.byref(m), C.byref(s), C.byref(v))
            res.append((e.value, m.value, s.value, v.value))
        return res

    def prepare_host_brownian(self) -> int:
        b = C.c_uint64()
        self.lib.check(self.lib.L.fmd_lmm_prepare_host_brownian(self.h, C.byref(b)))
        return b.value

    def step(self, vol_params=None, from_host=False) -> np.ndarray:
        """One simulation (n_periods Euler steps) + valuation of all calibration swaptions. Returns their values.
        from_host: False / 0 device-resident increments; True / 1 uploaded from pageable host doubles; 2 from pinned host
        doubles, asynchronously (fmc_vec_from_f64_pinned)."""
        out = np.empty(self.n_products)
        p = None
        if vol_params is not None:
            vp_ = np.ascontiguousarray(vol_params, dtype=np.float64)
            p = vp_.ctypes.data
        self.lib.check(self.lib.L.fmd_lmm_step(self.h, p, int(from_host), out.ctypes.data))
        return out

    def use_market_curve(self) -> None:
        """Initial forward rates from the EUR swap curve of LIBORMarketModelCalibrationATMTest.java:526-663 (instead of the synthetic curve)."""
        self.lib.check(self.lib.L.fmd_lmm_use_market_curve(self.h))
        self.n_products = self.lib.L.fmd_lmm_num_products(self.h)

    def forward_rates(self) -> np.ndarray:
        out = np.empty(self.n_periods)
        self.lib.L.fmd_lmm_get_forward_rates(self.h, out.ctypes.data)
        return out

    def set_parameters(self, params) -> None:
        p = np.ascontiguousarray(params, dtype=np.float64)
        assert p.size == self.n_parameters
        self.lib.L.fmd_lmm_set_parameters(self.h, p.ctypes.data)

    def calibrate(self, max_iterations: int = 200, accuracy: float = 1e-7, lmbda: float = 0.1, parameter_step: float = 1e-4) -> dict:
        """Levenberg-Marquardt calibration of the volatility parameters to the ATM swaption quotes with the settings of
        LIBORMarketModelCalibrationATMTest.java:317-340 by default. Every evaluation is one simulation + valuation of all products."""
        params = np.empty(self.n_parameters)
        info = np.empty(6)
        self.lib.check(self.lib.L.fmd_lmm_calibrate(self.h, int(max_iterations), float(accuracy), float(lmbda), float(parameter_step),
                                                    params.ctypes.data, info.ctypes.data))
        return {"parameters": params, "iterations": int(info[0]), "evaluations": int(info[1]), "rms_error": float(info[2]),
                "mean_deviation": float(info[3]), "seconds": float(info[4]), "seconds_per_evaluation": float(info[5])}

    def simulate(self) -> None:
        self.lib.check(self.lib.L.fmd_lmm_simulate(self.h))

    def set_valuation_threads(self, threads: int) -> None:
        """Host threads that value the calibration products of step() (the reference test uses one, T-ATM:319)."""
        self.lib.check(self.lib.L.fmd_lmm_set_valuation_threads(self.h, int(threads)))

    def set_price_products(self, on: bool) -> None:
        """How step() takes the averages: False (default) one blocking getAverage() per product while the next is recorded
        (implied-volatility products, T-ATM:261,511); True all product value vectors first, their averages in a second loop
        (price products, the objective function's own loop) — the runtime then batches the reductions."""
        self.lib.check(self.lib.L.fmd_lmm_set_price_products(self.h, 1 if on else 0))

    def libor(self, time_index: int, libor_index: int) -> np.ndarray:
        out = np.empty(self.local_paths)
        self.lib.check(self.lib.L.fmd_lmm_get_libor(self.h, time_index, libor_index, out.ctypes.data, out.size))
        return out

    def implied_vols(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.empty_like(v)
        self.lib.L.fmd_lmm_implied_vols(self.h, v.ctypes.data, out.ctypes.data)
        return out

    def bermudan(self, first_exercise: int, last_exercise: int, stride: int, swap_end: int, strike: float) -> float:
        v = C.c_double()
        self.lib.check(self.lib.L.fmd_lmm_bermudan(self.h, first_exercise, last_exercise, stride, swap_end, strike, C.byref(v)))
        return v.value
