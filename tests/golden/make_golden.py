#!/usr/bin/env python
"""Generates tests/golden/*.json: known-answer vectors that PIN the oracle (and through it the CUDA path).

Sources of truth, all independent of oracle/fm_oracle.c:
  * mt19937ar reference outputs (init_genrand(5489); init_by_array{0x123,0x234,0x345,0x456}) and numpy.random.MT19937
    (an independent MT19937 implementation) for the seeds the reference's tests use (31415, 314151, 1234, 53252);
  * scipy.special.ndtri (Cephes) for the inverse normal CDF;
  * the exact-arithmetic expectations of RandomVariableGPUTest.java:68-188 (computed here with numpy float32, which
    has the same IEEE binary32 semantics as Java float).
Run:  python tests/golden/make_golden.py      (numpy + scipy only; no reference checkout needed)
"""
import json
import os

import numpy as np
import scipy.special as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def mt_words(seed_key, count):
    rs = np.random.RandomState(seed_key)
    return [int(w) for w in rs._bit_generator.random_raw(count)]


def main():
    g = {}
    # --- MT19937 raw words -------------------------------------------------------------------------------------
    g["mt19937"] = {
        "init_genrand_5489_first3": [3499211612, 581869302, 3890346734],                              # mt19937ar.out
        "init_by_array_0x123_0x234_0x345_0x456_first5": [1067595299, 955945823, 477289528, 4107218783, 4228976476],
        "seeds_int": {str(s): mt_words(s, 16) for s in (31415, 314151, 1234, 53252, 5489)},              # setSeed(int)
        # setSeed(long s) == init_by_array{(int)(s >>> 32), (int)(s & 0xffffffff)}
        "seeds_long": {str(s): mt_words(np.array([0, s], dtype=np.uint32), 16) for s in (31415, 314151, 1234, 53252)},
        # deep offsets (words 1_000_000..1_000_007 and 40_000_000..40_000_003): pins the jump-ahead
        "seed_long_31415_at_1000000": None, "seed_long_31415_at_40000000": None,
    }
    rs = np.random.RandomState(np.array([0, 31415], dtype=np.uint32))
    w = rs._bit_generator.random_raw(40_000_004)
    g["mt19937"]["seed_long_31415_at_1000000"] = [int(x) for x in w[1_000_000:1_000_008]]
    g["mt19937"]["seed_long_31415_at_40000000"] = [int(x) for x in w[40_000_000:40_000_004]]
    # --- nextDouble + inverse normal -----------------------------------------------------------------------------
    w = np.array(g["mt19937"]["seeds_long"]["31415"], dtype=np.uint64)
    u = (((w[0::2] >> np.uint64(6)) << np.uint64(26)) | (w[1::2] >> np.uint64(6))).astype(np.float64) * 2.0 ** -52
    g["uniforms_seed_long_31415"] = [float(x) for x in u]
    ps = [1e-300, 1e-20, 1e-10, 1e-5, 0.001, 0.0749, 0.075, 0.0751, 0.3, 0.5, 0.7, 0.9249, 0.925, 0.9251, 0.999, 1 - 1e-10, 2.0 ** -52, 1 - 2.0 ** -52]
    g["icdf"] = {"p": ps, "ndtri": [float(sp.ndtri(p)) for p in ps], "rel_tol": 4e-15}
    g["brownian_seed_long_31415_dt0.5_first8"] = [float(np.float32(sp.ndtri(x) * np.sqrt(0.5))) for x in u]
    # --- RandomVariableGPUTest known answers ---------------------------------------------------------------------
    f = np.float32
    x = np.array([-4, -2, 0, 2, 4], dtype=np.float32)
    y = ((x + f(4.0)) / f(2.0) * f(2.0)) / f(2.0)
    g["rv_test"] = {
        "deterministic_chain": {"expect_average": 3.0, "expect_variance": 0.0},                # T-RV:68-86
        "stochastic_chain": {"input": [-4.0, -2.0, 0.0, 2.0, 4.0], "realizations": [float(v) for v in y], "average": 2.0, "variance": 2.0,
                             "times3_average": 6.0, "times3_variance": 18.0},              # T-RV:88-122
        "average_sizes": [2, 2, 3, 4, 5, 7, 10, 13, 99, 100, 1000, 1024, 2047, 2048, 2049, 20000, 200000],   # T-RV:127
        "sqrt_pow_input": [3.0, 1.0, 0.0, 2.0, 4.0, 1.0 / 3.0],                                   # T-RV:159
    }
    # --- Black-Scholes analytic (T-BS:146,156) -------------------------------------------------------------------
    g["black_scholes"] = {"S0": 1.0, "r": 0.05, "sigma": 0.30, "T": 2.0, "K": 1.05, "analytic": 0.18993678426215382, "tolerance": 0.005}
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(g, fh, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
