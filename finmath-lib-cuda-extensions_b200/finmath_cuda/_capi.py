"""ctypes binding of include/fmcuda.h (libfmcuda.so).

This is the Python twin of the JNI/FFM stub a Java maintainer would write. There is NO CPU
fallback: if the shared library is missing, or no CUDA device is present when a vector is first needed, the call
raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "lib", "libfmcuda.so"))

FMC_OK, FMC_ERR_INVALID, FMC_ERR_OOM, FMC_ERR_CUDA, FMC_ERR_SIZE, FMC_ERR_NOT_INIT, FMC_ERR_COMM, FMC_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7

# opcodes (include/fmcuda.h)
CAP, FLOOR, ADD, SUB, BUS, MULT, DIV, VID, POW = 1, 2, 3, 4, 5, 6, 7, 8, 9
SQUARED, SQRT, EXP, LOG, SIN, COS, INVERT, ABS, ISNAN = 20, 21, 22, 23, 24, 25, 26, 27, 28
ACCRUE, DISCOUNT, ADDPRODUCT, CHOOSE, ADDRATIO, SUBRATIO = 40, 41, 42, 43, 44, 45
RED_SUM, RED_AVERAGE, RED_VARIANCE, RED_SAMPLE_VARIANCE, RED_MIN, RED_MAX, RED_AVERAGE_W, RED_VARIANCE_W = 1, 2, 3, 4, 5, 6, 7, 8
UNIQUE_ID_BYTES = 128


class CudaError(RuntimeError):
    """JCuda's CudaException (RandomVariableCuda.java:167)."""


class FmcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes_in_use", "bytes_cached", "bytes_reserved", "bytes_high_water", "n_alloc", "n_alloc_reused",
        "n_ops_recorded", "n_kernels", "n_tape_kernels", "n_tape_instr", "n_nodes_stored", "n_nodes_fused",
        "n_flushes", "h2d_bytes", "d2h_bytes", "live_handles", "pending_nodes")]


# every symbol include/fmcuda.h declares: name -> (restype, argtypes)
_vec = C.c_uint64
_vecp = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)
PROTOTYPES = {
    "fmc_init": (C.c_int, [C.c_int]),
    "fmc_shutdown": (C.c_int, []),
    "fmc_is_initialized": (C.c_int, []),
    "fmc_last_error": (C.c_char_p, []),
    "fmc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fmc_device_info": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fmc_vec_from_f64": (C.c_int, [C.c_void_p, C.c_int64, _vecp]),
    "fmc_vec_from_f32": (C.c_int, [C.c_void_p, C.c_int64, _vecp]),
    "fmc_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "fmc_host_free": (C.c_int, [C.c_void_p]),
    "fmc_vec_from_f64_pinned": (C.c_int, [C.c_void_p, C.c_int64, _vecp]),
    "fmc_vec_fill": (C.c_int, [C.c_double, C.c_int64, _vecp]),
    "fmc_vec_alloc": (C.c_int, [C.c_int64, _vecp]),
    "fmc_vec_retain": (C.c_int, [_vec]),
    "fmc_vec_release": (C.c_int, [_vec]),
    "fmc_vec_size": (C.c_int, [_vec, C.POINTER(C.c_int64)]),
    "fmc_vec_to_f64": (C.c_int, [_vec, C.c_void_p, C.c_int64]),
    "fmc_vec_to_f32": (C.c_int, [_vec, C.c_void_p, C.c_int64]),
    "fmc_vec_get": (C.c_int, [_vec, C.c_int64, _f64p]),
    "fmc_vec_device_ptr": (C.c_int, [_vec, C.POINTER(C.c_void_p)]),
    "fmc_op_vs": (C.c_int, [C.c_int, _vec, C.c_double, _vecp]),
    "fmc_op_v": (C.c_int, [C.c_int, _vec, _vecp]),
    "fmc_op_vv": (C.c_int, [C.c_int, _vec, _vec, _vecp]),
    "fmc_op_vvs": (C.c_int, [C.c_int, _vec, _vec, C.c_double, _vecp]),
    "fmc_op_vvv": (C.c_int, [C.c_int, _vec, _vec, _vec, _vecp]),
    "fmc_op_choose": (C.c_int, [_vec, _vec, C.c_double, _vec, C.c_double, _vecp]),
    "fmc_reduce": (C.c_int, [C.c_int, _vec, _vec, _f64p]),
    "fmc_quantile": (C.c_int, [_vec, C.c_double, _f64p]),
    "fmc_quantile_expectation": (C.c_int, [_vec, C.c_double, C.c_double, _f64p]),
    "fmc_histogram": (C.c_int, [_vec, C.c_void_p, C.c_int, C.c_void_p]),
    "fmc_regression_normal_eq": (C.c_int, [_vecp, _f64p, C.c_int, _vec, C.c_void_p, C.c_void_p]),
    "fmc_brownian_generate": (C.c_int, [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, _vecp]),
    "fmc_mt19937_raw": (C.c_int, [C.c_int, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p]),
    "fmc_flush": (C.c_int, []),
    "fmc_sync": (C.c_int, []),
    "fmc_set_option": (C.c_int, [C.c_char_p, C.c_double]),
    "fmc_get_option": (C.c_int, [C.c_char_p, _f64p]),
    "fmc_get_stats": (C.c_int, [C.POINTER(FmcStats)]),
    "fmc_reset_stats": (C.c_int, []),
    "fmc_pool_trim": (C.c_int, []),
    "fmc_pool_purge": (C.c_int, []),
    "fmc_profile_read": (C.c_int, [_f64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "fmc_timer_start": (C.c_int, []),
    "fmc_timer_stop": (C.c_int, [C.POINTER(C.c_float)]),
    "fmc_comm_get_unique_id": (C.c_int, [C.c_char_p]),
    "fmc_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "fmc_comm_destroy": (C.c_int, []),
    "fmc_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libfmcuda.so (does not touch the GPU). Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback)")
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(status: int) -> None:
    if status == FMC_OK:
        return
    msg = (load().fmc_last_error() or b"").decode("utf-8", "replace")
    if status == FMC_ERR_OOM:
        raise MemoryError(msg)                      # java.lang.OutOfMemoryError, RandomVariableCuda.java:373-376
    if status == FMC_ERR_SIZE:
        raise IndexError(msg)                       # ArrayIndexOutOfBoundsException
    if status == FMC_ERR_INVALID:
        raise ValueError(msg)
    if status == FMC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)              # UnsupportedOperationException
    raise CudaError(f"[{status}] {msg}")


_initialized = False


def ensure_init(device_index: int | None = None) -> None:
    """Initialise the runtime on first use, like the static DeviceMemoryPool of RandomVariableCuda.java:562.

    The device comes from the same property the reference reads (RandomVariableCuda.java:161), taken from the
    environment: FINMATH_CUDA_DEVICE_INDEX (default: LOCAL_RANK if set by torchrun, else -1 = last device).
    """
    global _initialized
    if _initialized:
        return
    if device_index is None:
        env = os.environ.get("FINMATH_CUDA_DEVICE_INDEX")
        if env is not None:
            device_index = int(env)
        elif "LOCAL_RANK" in os.environ:
            device_index = int(os.environ["LOCAL_RANK"])
        elif int(os.environ.get("WORLD_SIZE", "1")) > 1:
            # several ranks, no device named: the default (-1, the last device) would put every rank on ONE GPU
            raise CudaError("WORLD_SIZE > 1 but neither LOCAL_RANK nor FINMATH_CUDA_DEVICE_INDEX is set: name the device of this rank")
        else:
            device_index = -1
    check(load().fmc_init(device_index))
    _initialized = True


def shutdown() -> None:
    global _initialized
    if _lib is not None:
        check(_lib.fmc_shutdown())
    _initialized = False


def stats() -> dict:
    s = FmcStats()
    check(load().fmc_get_stats(C.byref(s)))
    return {n: int(getattr(s, n)) for n, _ in FmcStats._fields_}


def profile_read() -> dict:
    ms, b, n = C.c_double(), C.c_uint64(), C.c_uint64()
    check(load().fmc_profile_read(C.byref(ms), C.byref(b), C.byref(n)))
    return {"tape_ms": ms.value, "tape_algorithmic_bytes": int(b.value), "tape_launches": int(n.value)}


def timer_start() -> None:
    check(load().fmc_timer_start())


def timer_stop() -> float:
    ms = C.c_float()
    check(load().fmc_timer_stop(C.byref(ms)))
    return float(ms.value)


def flush() -> None:
    """Execute every pending operation that is still referenced (no wait)."""
    check(load().fmc_flush())


def sync() -> None:
    """Execute everything pending and wait for the device (cuCtxSynchronize, RandomVariableCuda.java:472-476)."""
    check(load().fmc_sync())


def pool_trim() -> None:
    """RandomVariableCuda.clean() (RVC:751-753): cached device blocks go back to the driver."""
    check(load().fmc_pool_trim())


def reset_stats() -> None:
    check(load().fmc_reset_stats())


def set_option(key: str, value: float) -> None:
    check(load().fmc_set_option(key.encode(), float(value)))
