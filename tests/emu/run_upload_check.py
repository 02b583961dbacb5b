"""Host staging path on the tape-ISA emulator (tests/test_codegen_emulator.py): createRandomVariable(time, double[]) -> (float) cast on
the host worker pool -> getRealizations(), for ragged sizes around the block and chunk boundaries of the pool. The cast must be Java's
(float) (round to nearest even) whatever conversion loop (128- or 256-bit) and block hand-out the pool uses."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from finmath_cuda import _capi as capi  # noqa: E402

capi.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfmcuda_emu.so")
import finmath_cuda as fc  # noqa: E402

fc.ensure_init()
rng = np.random.default_rng(20261019)
for n in (1, 7, 1000, 32767, 32768, 65537, 262144 + 3, 1 << 20, (1 << 20) + 1, (1 << 22) + 12345):
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-46, 3.4028235677973366e38, 1.0 + 2.0 ** -24, 1.0 + 3 * 2.0 ** -24])
    idx = rng.integers(0, n, min(n, 64))
    x[idx] = special[np.arange(len(idx)) % len(special)]
    want = x.astype(np.float32).astype(np.float64)
    for rep in range(6):
        v = fc.RandomVariableCuda(0.0, x)
        got = np.asarray(v.getRealizations())
        assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), (n, rep)
        del v
print("upload / download round trips ok")
