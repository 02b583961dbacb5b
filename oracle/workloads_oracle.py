"""The workload drivers built on the CPU oracle backend (oracle/libfmdrivers_oracle.so). TEST INFRASTRUCTURE ONLY:
whole-workload parity checks and the CPU baseline of bench.py."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "finmath-lib-cuda-extensions_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from finmath_cuda.workloads import DriverLib  # noqa: E402  (generic ctypes binder only; no device code is touched)

ORACLE_DRIVER_LIB = os.path.join(_HERE, "libfmdrivers_oracle.so")


ORACLE_F64_DRIVER_LIB = os.path.join(_HERE, "libfmdrivers_oracle_f64.so")


def driver() -> DriverLib:
    """Drivers on the RandomVariableFromFloatArray twin (the parity oracle)."""
    return DriverLib(ORACLE_DRIVER_LIB)


def driver_f64() -> DriverLib:
    """Drivers on the RandomVariableFromDoubleArray twin (finmath-lib's default CPU type; timing baseline only)."""
    return DriverLib(ORACLE_F64_DRIVER_LIB)
