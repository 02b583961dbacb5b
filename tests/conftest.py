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
    finmath_cuda._capi.load()
    finmath_cuda.ensure_init()
    return finmath_cuda


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    oracle.lib()
    return oracle
