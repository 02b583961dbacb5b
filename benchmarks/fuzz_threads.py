#!/usr/bin/env python
"""Four host threads run fuzzed programs (tests/test_gpu_parity._random_program) concurrently against one runtime; every result is
checked against the oracle. usage: python benchmarks/fuzz_threads.py [emu]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "finmath-lib-cuda-extensions_b200"), ROOT+'/tests'):
    sys.path.insert(0, p)
import numpy as np
import finmath_cuda as fc
from finmath_cuda import _capi as capi
if len(sys.argv) > 1 and sys.argv[1] == 'emu':
    capi.LIB_PATH = ROOT + '/tests/emu/libfmcuda_emu.so'
capi.load(); fc.ensure_init()
from oracle import oracle as O
import test_gpu_parity as T
NT = 4; PER = 40
def inputs(seed):
    rs0 = np.random.default_rng(seed)
    n = int(rs0.choice([100, 513, 3000, 20000, 200000])); nops = int(rs0.choice([10, 40, 120, 400])); nleaf = int(rs0.integers(1, 6))
    xs = [np.random.default_rng(seed * 7 + i).uniform(0.2, 2.0, n) for i in range(nleaf)]
    return xs, nops
want = {}
for seed in range(NT * PER):
    xs, nops = inputs(seed)
    want[seed] = T._random_program(np.random.default_rng(seed), np.random.default_rng(seed + 99), [T._OracleRV(O, O.from_f64(x)) for x in xs], nops)
bad = []
def run(t):
    for seed in range(t * PER, (t + 1) * PER):
        xs, nops = inputs(seed)
        try:
            got = T._random_program(np.random.default_rng(seed), np.random.default_rng(seed + 99), [fc.RandomVariableCuda(0.0, x) for x in xs], nops)
            w = want[seed]
            ok = all(T.bits_equal(a, b) for a, b in zip(got[0], w[0])) and all(T._same_reduction(a, b) for a, b in zip(got[1], w[1]))
            if not ok: bad.append(("MISMATCH", seed))
        except Exception as e:
            bad.append(("ERROR", seed, str(e)[:150]))
t0 = time.time()
ths = [threading.Thread(target=run, args=(t,)) for t in range(NT)]
for th in ths: th.start()
for th in ths: th.join()
print("threads", NT, "programs", NT * PER, "failures", len(bad), bad[:5], "in", round(time.time() - t0, 1), "s")
