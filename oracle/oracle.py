"""ctypes loader for the CPU oracle (oracle/fm_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under finmath-lib-cuda-extensions_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfm_oracle.so")

# opcodes (fm_oracle.h)
CAP, FLOOR, ADD, SUB, BUS, MULT, DIV, VID, POW = 1, 2, 3, 4, 5, 6, 7, 8, 9
SQUARED, SQRT, EXP, LOG, SIN, COS, INVERT, ABS, ISNAN = 20, 21, 22, 23, 24, 25, 26, 27, 28
ACCRUE, DISCOUNT, ADDPRODUCT, CHOOSE, ADDRATIO, SUBRATIO = 40, 41, 42, 43, 44, 45


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fm_oracle.c")
    hdr = os.path.join(_HERE, "fm_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    f32p, f64p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    i64, dbl, i32 = C.c_int64, C.c_double, C.c_int
    L.orc_from_f64.argtypes = [f64p, f32p, i64]
    L.orc_op_vs.argtypes = [i32, f32p, dbl, f32p, i64]
    L.orc_op_v.argtypes = [i32, f32p, f32p, i64]
    L.orc_op_vv.argtypes = [i32, f32p, f32p, f32p, i64]
    L.orc_op_vvs.argtypes = [i32, f32p, f32p, dbl, f32p, i64]
    L.orc_op_vvv.argtypes = [i32, f32p, f32p, f32p, f32p, i64]
    for name in ("orc_min", "orc_max", "orc_average", "orc_variance", "orc_sample_variance"):
        getattr(L, name).argtypes = [f32p, i64]
        getattr(L, name).restype = dbl
    for name in ("orc_average_w", "orc_variance_w"):
        getattr(L, name).argtypes = [f32p, f32p, i64]
        getattr(L, name).restype = dbl
    L.orc_average_f64.argtypes = [f64p, i64]
    L.orc_average_f64.restype = dbl
    L.orc_quantile.argtypes = [f32p, i64, dbl]
    L.orc_quantile.restype = dbl
    L.orc_quantile_expectation.argtypes = [f32p, i64, dbl, dbl]
    L.orc_quantile_expectation.restype = dbl
    L.orc_histogram.argtypes = [f32p, i64, f64p, i32, f64p]
    L.orc_mt_fill_u32.argtypes = [i32, i64, C.c_uint64, u32p, i64]
    L.orc_icdf.argtypes = [dbl]
    L.orc_icdf.restype = dbl
    L.orc_icdf_array.argtypes = [f64p, f64p, i64]
    L.orc_brownian.argtypes = [i32, i64, i32, i32, i64, i64, f64p, f32p]
    L.orc_brownian_f64.argtypes = [i32, i64, i32, i32, i64, i64, f64p, f64p]
    L.orc_regression_normal_eq.argtypes = [C.POINTER(f32p), f64p, i32, f32p, i64, f64p, f64p]
    _lib = L
    return L


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p32(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _p64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def from_f64(values) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.float64)
    out = np.empty(v.shape, dtype=np.float32)
    lib().orc_from_f64(_p64(v), _p32(out), v.size)
    return out


def op_vs(op: int, x, s: float) -> np.ndarray:
    x = _f32(x); out = np.empty_like(x)
    assert lib().orc_op_vs(op, _p32(x), float(s), _p32(out), x.size) == 0
    return out


def op_v(op: int, x) -> np.ndarray:
    x = _f32(x); out = np.empty_like(x)
    assert lib().orc_op_v(op, _p32(x), _p32(out), x.size) == 0
    return out


def op_vv(op: int, x, y) -> np.ndarray:
    x = _f32(x); y = _f32(y); out = np.empty_like(x)
    assert x.size == y.size
    assert lib().orc_op_vv(op, _p32(x), _p32(y), _p32(out), x.size) == 0
    return out


def op_vvs(op: int, x, y, s: float) -> np.ndarray:
    x = _f32(x); y = _f32(y); out = np.empty_like(x)
    assert lib().orc_op_vvs(op, _p32(x), _p32(y), float(s), _p32(out), x.size) == 0
    return out


def op_vvv(op: int, x, y, z) -> np.ndarray:
    x = _f32(x); y = _f32(y); z = _f32(z); out = np.empty_like(x)
    assert lib().orc_op_vvv(op, _p32(x), _p32(y), _p32(z), _p32(out), x.size) == 0
    return out


def minimum(x) -> float:
    x = _f32(x); return lib().orc_min(_p32(x), x.size)


def maximum(x) -> float:
    x = _f32(x); return lib().orc_max(_p32(x), x.size)


def average(x, prob=None) -> float:
    x = _f32(x)
    if prob is None:
        return lib().orc_average(_p32(x), x.size)
    p = _f32(prob)
    return lib().orc_average_w(_p32(x), _p32(p), x.size)


def variance(x, prob=None) -> float:
    x = _f32(x)
    if prob is None:
        return lib().orc_variance(_p32(x), x.size)
    p = _f32(prob)
    return lib().orc_variance_w(_p32(x), _p32(p), x.size)


def sample_variance(x) -> float:
    x = _f32(x); return lib().orc_sample_variance(_p32(x), x.size)


def quantile(x, q: float) -> float:
    x = _f32(x); return lib().orc_quantile(_p32(x), x.size, float(q))


def quantile_expectation(x, q0: float, q1: float) -> float:
    x = _f32(x); return lib().orc_quantile_expectation(_p32(x), x.size, float(q0), float(q1))


def histogram(x, interval_points) -> np.ndarray:
    x = _f32(x)
    pts = np.ascontiguousarray(interval_points, dtype=np.float64)
    out = np.empty(pts.size + 1, dtype=np.float64)
    lib().orc_histogram(_p32(x), x.size, _p64(pts), pts.size, _p64(out))
    return out


SEED_LONG, SEED_INT = 0, 1


def mt_u32(seed: int, count: int, seed_mode: int = SEED_LONG, skip: int = 0) -> np.ndarray:
    out = np.empty(count, dtype=np.uint32)
    lib().orc_mt_fill_u32(seed_mode, seed, skip, out.ctypes.data_as(C.POINTER(C.c_uint32)), count)
    return out


def mt_doubles_from_u32(words: np.ndarray) -> np.ndarray:
    """BitsStreamGenerator.nextDouble on consecutive word pairs."""
    w = words.astype(np.uint64)
    hi = (w[0::2] >> np.uint64(6)) << np.uint64(26)
    lo = w[1::2] >> np.uint64(6)
    return (hi | lo).astype(np.float64) * 2.0 ** -52


def icdf(p) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    lib().orc_icdf_array(_p64(p), _p64(out), p.size)
    return out


def brownian(seed: int, T: int, F: int, n: int, sqrt_dt, seed_mode: int = SEED_LONG, p0: int = 0, p1: int | None = None,
             dtype=np.float32) -> np.ndarray:
    """Returns array [T*F, p1-p0] (row t*F+f)."""
    p1 = n if p1 is None else p1
    sd = np.ascontiguousarray(sqrt_dt, dtype=np.float64)
    assert sd.size == T
    out = np.empty((T * F, p1 - p0), dtype=dtype)
    if dtype == np.float32:
        lib().orc_brownian(seed_mode, seed, T, F, p0, p1, _p64(sd), _p32(out))
    else:
        lib().orc_brownian_f64(seed_mode, seed, T, F, p0, p1, _p64(sd), _p64(out))
    return out


def regression_normal_eq(basis, y):
    """basis: list of float32 arrays or python floats (deterministic). Returns (XtX[k,k], Xty[k])."""
    k = len(basis)
    y = _f32(y)
    arrs = [None if np.isscalar(b) else _f32(b) for b in basis]
    ptrs = (C.POINTER(C.c_float) * k)(*[(_p32(a) if a is not None else C.POINTER(C.c_float)()) for a in arrs])
    scal = np.array([float(b) if np.isscalar(b) else 0.0 for b in basis], dtype=np.float64)
    XtX = np.empty((k, k), dtype=np.float64)
    Xty = np.empty(k, dtype=np.float64)
    lib().orc_regression_normal_eq(ptrs, _p64(scal), k, _p32(y), y.size, _p64(XtX), _p64(Xty))
    return XtX, Xty
