"""benchmarks/window_check.py on the tape-ISA emulator (tests/test_codegen_emulator.py): the LMM driver records its simulation through
the product's host code, the emulator executes the tapes (the Brownian increments are the emulator's stand-in values, which is all
a comparison between two schedules of the same arithmetic needs). Run with LD_PRELOAD=libfmcuda_emu.so FMC_EMU_FAKE_BROWNIAN=1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
from finmath_cuda import _capi as capi  # noqa: E402

capi.LIB_PATH = os.path.join(ROOT, "tests", "emu", "libfmcuda_emu.so")
sys.argv = ["window_check.py"] + sys.argv[1:]
exec(compile(open(os.path.join(ROOT, "benchmarks", "window_check.py")).read(), "window_check.py", "exec"))
