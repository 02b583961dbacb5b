"""Runtime life cycle on the GPU: fmc_shutdown releases everything (device ring of long tapes, copy stream, result slots,
cached tapes, jump-ahead states) and a second fmc_init starts from scratch with identical results. Runs in a child process
because a shutdown invalidates every handle the other test modules hold."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys
sys.path.insert(0, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200")); sys.path.insert(0, ROOT)
import numpy as np
import finmath_cuda as fc

def work():
    rng = np.random.default_rng(3)
    x = fc.RandomVariableCuda(0.0, rng.uniform(0.5, 1.5, 70001))
    y = x
    for k in range(700):                       # a tape too long for the inline argument block
        y = y.mult(1.0 + 1e-4 * k).add(x).discount(x, 0.25)
    bm = fc.BrownianMotionCuda(fc.TimeDiscretization(0.0, 5, 0.5), 1, 40000, 31415)
    inc = bm.getBrownianIncrement(3, 0)
    return y.getAverage(), float(y.getRealizationsFloat()[12345]), inc.getVariance(), fc.stats()["live_handles"]

fc.ensure_init()
a = work()
fc.shutdown()
assert fc._capi.load().fmc_is_initialized() == 0
fc.ensure_init()
b = work()
s = fc.stats()
fc.shutdown()
assert a[:3] == b[:3], (a, b)
print("LIFECYCLE OK", a[0], s["bytes_in_use"])
'''


def test_shutdown_and_reinit_in_a_fresh_process():
    env = dict(os.environ)
    env.pop("FMC_TEST_TAPE_EMULATOR", None)
    r = subprocess.run([sys.executable, "-c", f"ROOT = {ROOT!r}\n" + CHILD], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "LIFECYCLE OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
