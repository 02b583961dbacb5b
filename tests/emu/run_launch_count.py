"""Launches per LMM step on the tape-ISA emulator (tests/test_codegen_emulator.py): the same simulation + 144 valuations recorded again
and again must be cut into the same windows every time. (A count of still-referenced targets that included recycled node slots twice
once sent every third step's simulation through one cone per flush: 213 instead of 165 launches, 10 % more kernel time on the GPU.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from finmath_cuda import _capi as capi  # noqa: E402

capi.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfmcuda_emu.so")
import finmath_cuda as fc  # noqa: E402
from finmath_cuda.workloads import DriverLib  # noqa: E402

fc.ensure_init()
paths = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = DriverLib().lmm(paths, 80, 0.5, 1, 31415, 0, (0, paths))
counts = []
for i in range(10):
    k0 = fc.stats()["n_tape_kernels"]
    m.step()
    counts.append(fc.stats()["n_tape_kernels"] - k0)
m.close()
print("tape launches per step:", counts)
# 144 valuations + the windows of 80 time steps (27 at three levels per window) + a few ragged ones
ok = max(counts) <= min(counts) + 8 and max(counts) <= 144 + 40
print("launch count stable", ok)
sys.exit(0 if ok else 1)
