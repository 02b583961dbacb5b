import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fc():
    """The product package, initialised on cuda:0. GPU tests fail loudly if the native library is missing."""
    import finmath_cuda
    emu = os.environ.get("FMC_TEST_TAPE_EMULATOR")
    if emu:
        # tests/test_codegen_emulator.py only: the product's host code linked against the tape-ISA emulator
        # (tests/emu/), to check the code generator on a machine without a GPU. Never set on the GPU box.
        finmath_cuda._capi.LIB_PATH = emu
    finmath_cuda._capi.load()
    finmath_cuda.ensure_init()
    for kv in filter(None, os.environ.get("FMC_TEST_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        finmath_cuda.set_option(k, float(v))
    return finmath_cuda


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle
