"""CPU tests (no GPU): the oracle against the golden vectors / known answers that pin it."""
import json
import math
import os

import numpy as np
import pytest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_mt19937_known_answers(O):
    g = GOLDEN["mt19937"]
    assert list(O.mt_u32(5489, 3, O.SEED_INT)) == g["init_genrand_5489_first3"]
    for seed, words in g["seeds_int"].items():
        assert list(O.mt_u32(int(seed), len(words), O.SEED_INT)) == words
    for seed, words in g["seeds_long"].items():
        assert list(O.mt_u32(int(seed), len(words), O.SEED_LONG)) == words
    assert list(O.mt_u32(31415, 8, O.SEED_LONG, skip=1_000_000)) == g["seed_long_31415_at_1000000"]


def test_mt19937_init_by_array_reference_vector(O):
    import ctypes as C
    L = O.lib()

    class MT(C.Structure):
        _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int)]
    g = MT()
    key = (C.c_uint32 * 4)(0x123, 0x234, 0x345, 0x456)
    L.orc_mt_seed_array(C.byref(g), key, 4)
    L.orc_mt_next_u32.restype = C.c_uint32
    got = [L.orc_mt_next_u32(C.byref(g)) for _ in range(5)]
    assert got == GOLDEN["mt19937"]["init_by_array_0x123_0x234_0x345_0x456_first5"]


def test_mt19937_matches_numpy_for_long_streams(O):
    rs = np.random.RandomState(np.array([0, 53252], dtype=np.uint32))
    assert np.array_equal(O.mt_u32(53252, 300_000, O.SEED_LONG), rs._bit_generator.random_raw(300_000).astype(np.uint32))


def test_next_double_and_icdf(O):
    u = O.mt_doubles_from_u32(O.mt_u32(31415, 16, O.SEED_LONG))
    assert list(u) == GOLDEN["uniforms_seed_long_31415"]
    g = GOLDEN["icdf"]
    got = O.icdf(np.array(g["p"]))
    for a, b in zip(got, g["ndtri"]):
        assert abs(a - b) <= g["rel_tol"] * max(abs(b), 1.0), (a, b)
    inc = O.brownian(31415, 1, 1, 8, [math.sqrt(0.5)])
    assert [float(v) for v in inc[0]] == GOLDEN["brownian_seed_long_31415_dt0.5_first8"]


def test_brownian_layout_and_slices(O):
    T, F, n = 5, 3, 40
    sq = np.sqrt(np.full(T, 0.25))
    full = O.brownian(1234, T, F, n, sq)
    # path-major draw order: element (t,f) of path p uses words 2*((p*T+t)*F+f), +1
    w = O.mt_u32(1234, 2 * T * F * n)
    u = O.mt_doubles_from_u32(w).reshape(n, T, F)
    want = (O.icdf(u) * 0.5).astype(np.float32)
    assert np.array_equal(full.reshape(T, F, n), want.transpose(1, 2, 0))
    parts = [O.brownian(1234, T, F, n, sq, p0=a, p1=b) for a, b in ((0, 12), (12, 13), (13, 40))]
    assert np.array_equal(np.concatenate(parts, axis=1), full)


def test_random_variable_test_known_answers(O):
    g = GOLDEN["rv_test"]
    x = O.from_f64(g["stochastic_chain"]["input"])
    y = O.op_vs(O.DIV, O.op_vs(O.MULT, O.op_vs(O.DIV, O.op_vs(O.ADD, x, 4.0), 2.0), 2.0), 2.0)
    assert [float(v) for v in y] == g["stochastic_chain"]["realizations"]
    assert O.average(y) == 2.0 and O.variance(y) == 2.0                     # T-RV:106,111 (exact equality)
    y3 = O.op_vs(O.MULT, y, 3.0)
    assert O.average(y3) == 6.0 and O.variance(y3) == 18.0                  # T-RV:118,121
    for size in g["average_sizes"]:                                         # T-RV:124-153
        v = np.arange(size, dtype=np.float64)
        assert abs(O.average(O.from_f64(v)) - (size - 1) / 2.0) <= (size - 1) / 2.0 * 1e-6 + 1e-300
        want = 0.5 if size % 2 == 0 else (size // 2) / size
        assert abs(O.average(O.from_f64(v % 2)) - want) <= size / 2.0 * 1e-7
    z = O.from_f64(g["sqrt_pow_input"])                                     # T-RV:155-188
    d = O.op_vv(O.SUB, O.op_v(O.SQRT, z), O.op_vs(O.POW, z, 0.5))
    assert abs(O.average(d)) <= 1e-7 and O.variance(d) <= 1e-7
    d = O.op_vv(O.SUB, O.op_v(O.SQUARED, z), O.op_vs(O.POW, z, 2.0))
    assert abs(O.average(d)) <= 1e-7 and O.variance(d) <= 1e-7


def test_java_float_semantics(O):
    f = np.float32
    x = np.array([0.1, 1 / 3, 2.5e-8, 7.0], dtype=np.float32)
    y = np.array([0.7, 3.0, 1e8, 1 / 3], dtype=np.float32)
    # no FMA: x + y*s rounds the product first (RVF:1345-1348)
    s = 1.0 / 3.0
    want = x + (y * f(s))
    assert np.array_equal(O.op_vvs(O.ADDPRODUCT, x, y, s), want)
    assert np.array_equal(O.op_vvs(O.ACCRUE, x, y, s), x * (f(1.0) + y * f(s)))
    assert np.array_equal(O.op_vvs(O.DISCOUNT, x, y, s), x / (f(1.0) + y * f(s)))
    # double-then-round transcendentals (RVF:849,890,905,920)
    assert np.array_equal(O.op_v(O.EXP, x), np.exp(x.astype(np.float64)).astype(np.float32))
    assert np.array_equal(O.op_vs(O.POW, x, 0.3), np.power(x.astype(np.float64), float(f(0.3))).astype(np.float32))
    # Math.min / Math.max: NaN propagating, -0 < +0 (RVF:759,774)
    a = np.array([0.0, -0.0, np.nan, 1.0], dtype=np.float32); b = np.array([-0.0, 0.0, 1.0, np.nan], dtype=np.float32)
    mn, mx = O.op_vv(O.CAP, a, b), O.op_vv(O.FLOOR, a, b)
    assert np.signbit(mn[0]) and np.signbit(mn[1]) and np.isnan(mn[2]) and np.isnan(mn[3])
    assert not np.signbit(mx[0]) and not np.signbit(mx[1]) and np.isnan(mx[2]) and np.isnan(mx[3])
    # choose: x >= 0 ? a : b, NaN trigger selects b (RVF:1281)
    t = np.array([0.0, -0.0, np.nan, -1.0], dtype=np.float32)
    assert list(O.op_vvv(O.CHOOSE, t, np.ones(4, np.float32), np.zeros(4, np.float32))) == [1.0, 1.0, 0.0, 0.0]
    # Math.pow corner cases
    assert np.isnan(O.op_vs(O.POW, np.array([1.0], np.float32), float("nan"))[0])
    assert O.op_vs(O.POW, np.array([np.nan], np.float32), 0.0)[0] == 1.0


def test_reductions_definitions(O):
    rng = np.random.RandomState(7)
    x = rng.standard_normal(10007).astype(np.float32)
    p = rng.random_sample(10007).astype(np.float32)
    xd = x.astype(np.float64)
    assert abs(O.average(x) - math.fsum(xd) / x.size) < 1e-15
    avg = O.average(x)
    assert abs(O.variance(x) - math.fsum((xd - avg) ** 2) / x.size) < 1e-14              # RVF:360-382 (biased)
    assert abs(O.sample_variance(x) - O.variance(x) * x.size / (x.size - 1)) < 1e-15     # RVF:418
    assert abs(O.average(x, p) - math.fsum(xd * p.astype(np.float64)) / x.size) < 1e-15  # RVF:337-357
    aw = O.average(x, p)
    assert abs(O.variance(x, p) - math.fsum((xd - aw) ** 2 * p.astype(np.float64))) < 1e-10   # RVF:385-407: NOT / n
    assert O.minimum(x) == float(x.min()) and O.maximum(x) == float(x.max())
    assert math.isnan(O.average(np.zeros(0, np.float32))) and O.variance(np.zeros(1, np.float32)) == 0.0
    s = np.sort(x)
    for q in (0.0, 0.25, 0.5, 0.99, 1.0):                                                 # RVF:484
        idx = min(max(int(math.floor((x.size + 1) * q - 1 + 0.5)), 0), x.size - 1)
        assert O.quantile(x, q) == float(s[idx])
    h = O.histogram(x, [-1.0, 0.0, 1.0])
    assert abs(h.sum() - 1.0) < 1e-15 and abs(h[0] - (x <= -1.0).mean()) < 1e-15


def test_regression_normal_equations_definition(O):
    rng = np.random.RandomState(3)
    n = 5003
    b1, b2, y = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    XtX, Xty = O.regression_normal_eq([1.0, b1, b2], y)
    assert XtX[0, 0] == 1.0
    assert abs(XtX[0, 1] - O.average(O.op_vs(O.MULT, b1, 1.0))) == 0.0
    assert XtX[1, 2] == XtX[2, 1] == O.average(O.op_vv(O.MULT, b1, b2))
    assert Xty[2] == O.average(O.op_vv(O.MULT, y, b2))


def test_workload_drivers_on_oracle_backend():
    """Config 1 of BASELINE.json on the CPU path: Black-Scholes 100k paths x 100 steps within 0.005 of analytic (T-BS:156)."""
    from oracle.workloads_oracle import driver
    d = driver()
    v, a = d.bs_call(100_000)
    assert abs(a - GOLDEN["black_scholes"]["analytic"]) < 1e-12
    assert abs(v - a) < GOLDEN["black_scholes"]["tolerance"]
    m = d.lmm(512)
    assert m.n_products == 144 and m.n_parameters == 48        # T-ATM: 154 products, 10 beyond the 40y grid
    vals = m.step()
    iv = m.implied_vols(vals)
    assert np.all(np.isfinite(vals)) and np.all(vals > 0) and np.all(np.abs(iv - 0.005) < 0.002)
    assert m.bermudan(10, 30, 2, 40, 0.02) > 0


def test_double_array_twin_agrees_with_the_float_twin():
    """oracle/RandomVariableFromDoubleArray.hpp (finmath-lib's default CPU type, CPU timing baseline only) runs the same
    drivers; with every intermediate in double it must agree with the float twin to float accuracy."""
    from oracle.workloads_oracle import driver, driver_f64
    f32, f64 = driver(), driver_f64()
    (v32, a32), (v64, a64) = f32.bs_call(100_000), f64.bs_call(100_000)
    assert a32 == a64 and abs(v32 - v64) < 1e-6 * abs(v64)
    m32, m64 = f32.lmm(2000), f64.lmm(2000)
    s32, s64 = m32.step(), m64.step()
    assert np.allclose(s32, s64, rtol=2e-4, atol=1e-9)
    m32.simulate(); m64.simulate()
    b32, b64 = m32.bermudan(10, 30, 2, 40, 0.02), m64.bermudan(10, 30, 2, 40, 0.02)
    assert abs(b32 - b64) < 5e-3 * abs(b64)         # exercise decisions of a few paths flip between float and double
