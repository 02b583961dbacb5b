"""GPU parity tests proper: the CUDA path (through the C ABI / host mirror) against the CPU oracle on the same
seeded inputs. Bit-exact for + - * / min max abs choose and the uint32 random stream; <= 1 float ulp for the
double-then-round transcendentals; 1e-5 relative for reductions and regression (north-star tolerances)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 100_000          # RandomVariableGPUTest.java:196


def ulp_diff(a: np.ndarray, b: np.ndarray) -> int:
    a = np.asarray(a, dtype=np.float32); b = np.asarray(b, dtype=np.float32)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    ia = a.view(np.int32).astype(np.int64); ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    d = np.abs(ia - ib)
    d[nan_a] = 0
    return int(d.max()) if d.size else 0


def bits_equal(a, b) -> bool:
    a = np.asarray(a, dtype=np.float32); b = np.asarray(b, dtype=np.float32)
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | nan))


@pytest.fixture(scope="module")
def data(O):
    # inputs: commons-math3 style uniforms from the MT stream the generator under test also produces
    x = O.mt_doubles_from_u32(O.mt_u32(31415, 2 * N, O.SEED_INT))
    y = O.mt_doubles_from_u32(O.mt_u32(314151, 2 * N, O.SEED_INT))
    z = O.mt_doubles_from_u32(O.mt_u32(1234, 2 * N, O.SEED_INT)) - 0.5
    return x, y, z


def test_upload_download_roundtrip(fc, O, data):
    x, _, _ = data
    rv = fc.RandomVariableCuda(0.0, x)
    assert rv.size() == N and not rv.isDeterministic() and rv.getTypePriority() == 20
    assert bits_equal(rv.getRealizationsFloat(), O.from_f64(x))
    assert np.array_equal(rv.getRealizations(), O.from_f64(x).astype(np.float64))
    assert rv.get(17) == float(np.float32(x[17]))


SCALARS = [1.0 / 3.0, 3.1415, 2.0, -0.25, 0.0]


@pytest.mark.parametrize("name", ["cap", "floor", "add", "sub", "bus", "mult", "div", "vid"])
def test_scalar_ops_bit_exact(fc, O, data, name):
    x, _, z = data
    for src in (x, z):
        rv = fc.RandomVariableCuda(0.0, src)
        xf = O.from_f64(src)
        for s in SCALARS:
            got = getattr(rv, name)(s).getRealizationsFloat()
            want = O.op_vs(getattr(O, name.upper()), xf, s)
            assert bits_equal(got, want), (name, s)


@pytest.mark.parametrize("name", ["add", "sub", "bus", "mult", "div", "vid", "cap", "floor"])
def test_vector_ops_bit_exact(fc, O, data, name):
    x, y, z = data
    a, b = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, z)
    got = getattr(a, name)(b).getRealizationsFloat()
    want = O.op_vv(getattr(O, name.upper()), O.from_f64(x), O.from_f64(z))
    assert bits_equal(got, want)
    # x op x (same operand twice)
    got = getattr(a, name)(a).getRealizationsFloat()
    want = O.op_vv(getattr(O, name.upper()), O.from_f64(x), O.from_f64(x))
    assert bits_equal(got, want)


@pytest.mark.parametrize("name,exact", [("squared", True), ("sqrt", True), ("invert", True), ("abs", True), ("isNaN", True),
                                        ("exp", False), ("log", False), ("sin", False), ("cos", False)])
def test_unary_ops(fc, O, data, name, exact):
    x, _, z = data
    for src in (x, z * 20.0):
        rv = fc.RandomVariableCuda(0.0, src)
        got = getattr(rv, name)().getRealizationsFloat()
        want = O.op_v(getattr(O, name.upper()), O.from_f64(src))
        if exact:
            assert bits_equal(got, want), name
        else:
            assert ulp_diff(got, want) <= 1, (name, ulp_diff(got, want))


@pytest.mark.parametrize("name", ["exp", "log"])
def test_exp_log_bit_equal_to_the_oracle_on_a_dense_sample_of_all_floats(fc, O, name):
    """exp / log are range reduction + a double-precision FMA polynomial written out in the interpreter, rounded once to float
    (RVF:903-921: (float)Math.exp((double)x)). The same sequence agrees with glibc on ALL 2^32 inputs on the CPU
    (benchmarks/micro/explog_exhaustive.c); here every 256th float bit pattern (plus the neighbourhood of the special values)
    goes through the GPU and must be bit-equal to the oracle; benchmarks/explog_gpu_exhaustive.py runs all 2^32."""
    bits = np.arange(0, 1 << 32, 256, dtype=np.uint64).astype(np.uint32)
    special = np.array([0x00000000, 0x80000000, 0x00000001, 0x007fffff, 0x00800000, 0x3f800000, 0x3f7fffff, 0x3f800001, 0x7f7fffff, 0x7f800000,
                        0xff800000, 0x7fc00000, 0x42b17218, 0x42b17217, 0x42b17219, 0xc2cff1b5, 0xc2cff1b4, 0xc2aeac50, 0xc2aeac4f], dtype=np.uint32)
    bits = np.concatenate([bits, special, special + np.uint32(1), special - np.uint32(1)])
    x = bits.view(np.float32)
    X = fc.RandomVariableCuda(0.0, x)
    got = getattr(X, name)().getRealizationsFloat()
    want = O.op_v(O.EXP if name == "exp" else O.LOG, x)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    bad = np.flatnonzero(~same)
    assert bad.size == 0, (name, bad.size, [(hex(int(bits[i])), float(got[i]), float(want[i])) for i in bad[:5]])


@pytest.mark.parametrize("e", [2.0, 0.5, 3.0, -1.5, 1.0 / 3.0, 0.0, 1.0])
def test_pow(fc, O, data, e):
    x, _, _ = data
    rv = fc.RandomVariableCuda(0.0, x * 4.0)
    got = rv.pow(e).getRealizationsFloat()
    want = O.op_vs(O.POW, O.from_f64(x * 4.0), e)
    assert ulp_diff(got, want) <= 1, ulp_diff(got, want)


def test_compound_ops_bit_exact(fc, O, data):
    x, y, z = data
    X, Y, Z = (fc.RandomVariableCuda(0.0, v) for v in (x, y, z))
    xf, yf, zf = O.from_f64(x), O.from_f64(y), O.from_f64(z)
    for p in (2.0, 1.0 / 3.0):
        assert bits_equal(X.accrue(Y, p).getRealizationsFloat(), O.op_vvs(O.ACCRUE, xf, yf, p))
        assert bits_equal(X.discount(Y, p).getRealizationsFloat(), O.op_vvs(O.DISCOUNT, xf, yf, p))
        assert bits_equal(X.addProduct(Y, p).getRealizationsFloat(), O.op_vvs(O.ADDPRODUCT, xf, yf, p))
    assert bits_equal(X.addProduct(Y, Z).getRealizationsFloat(), O.op_vvv(O.ADDPRODUCT, xf, yf, zf))
    assert bits_equal(X.addRatio(Z, Y).getRealizationsFloat(), O.op_vvv(O.ADDRATIO, xf, zf, yf))
    assert bits_equal(X.subRatio(Z, Y).getRealizationsFloat(), O.op_vvv(O.SUBRATIO, xf, zf, yf))
    assert bits_equal(Z.choose(X, Y).getRealizationsFloat(), O.op_vvv(O.CHOOSE, zf, xf, yf))
    # choose with deterministic branches
    got = Z.choose(fc.RandomVariableCuda(1.5), Y).getRealizationsFloat()
    assert bits_equal(got, O.op_vvv(O.CHOOSE, zf, np.full(N, 1.5, dtype=np.float32), yf))
    got = Z.choose(X, fc.RandomVariableCuda(-2.0)).getRealizationsFloat()
    assert bits_equal(got, O.op_vvv(O.CHOOSE, zf, xf, np.full(N, -2.0, dtype=np.float32)))
    # addSumProduct = fold of addProduct (RVF:1385-1392)
    got = X.addSumProduct([X, Y], [Y, Z]).getRealizationsFloat()
    want = O.op_vvv(O.ADDPRODUCT, O.op_vvv(O.ADDPRODUCT, xf, xf, yf), yf, zf)
    assert bits_equal(got, want)


def test_edge_values_min_max_nan_zero(fc, O):
    nan, inf = float("nan"), float("inf")
    a = np.array([0.0, -0.0, 0.0, -0.0, nan, 1.0, nan, inf, -inf, 1e-45, 3.0, -1.0], dtype=np.float32)
    b = np.array([-0.0, 0.0, 0.0, -0.0, 1.0, nan, nan, -inf, inf, -1e-45, 3.0, 2.0], dtype=np.float32)
    A, B = fc.RandomVariableCuda(0.0, a), fc.RandomVariableCuda(0.0, b)
    for name in ("cap", "floor", "add", "sub", "mult", "div", "vid"):
        got = getattr(A, name)(B).getRealizationsFloat()
        want = O.op_vv(getattr(O, name.upper()), a, b)
        assert bits_equal(got, want), (name, got, want)
    for s in (0.0, -0.0, nan, 1.0):
        assert bits_equal(A.cap(s).getRealizationsFloat(), O.op_vs(O.CAP, a, s)), s
        assert bits_equal(A.floor(s).getRealizationsFloat(), O.op_vs(O.FLOOR, a, s)), s
    assert bits_equal(A.isNaN().getRealizationsFloat(), O.op_v(O.ISNAN, a))
    assert bits_equal(A.abs().getRealizationsFloat(), O.op_v(O.ABS, a))
    assert bits_equal(A.choose(B, A).getRealizationsFloat(), O.op_vvv(O.CHOOSE, a, b, a))
    for e in (nan, 0.0, inf, 2.0, 0.5):
        assert ulp_diff(A.pow(e).getRealizationsFloat(), O.op_vs(O.POW, a, e)) <= 1, e


def test_division_exact_over_the_whole_float_range(fc, O):
    """The interpreter's division is a hand-scheduled Markstein sequence with a conservative range check and an
    out-of-line div.rn.f32 for everything else: exercise both sides of the check and the boundary, per lane and mixed
    within a warp — random bit patterns (all exponents, denormals, NaN, inf), exact powers of two around 2^-57 / 2^58,
    signed zeros, and quotients that under/overflow. div, vid, discount and both scalar forms must be bit exact."""
    rng = np.random.default_rng(99)
    n = 1 << 16
    bits_a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    bits_b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    a = bits_a.view(np.float32).copy(); b = bits_b.view(np.float32).copy()
    # boundary exponents and special values sprinkled through otherwise ordinary data
    special = np.array([0.0, -0.0, 1.0, -1.0, 2.0 ** -57, 2.0 ** -58, 2.0 ** 57, 2.0 ** 58, np.nextafter(np.float32(2.0 ** 58), np.float32(0)),
                        1e-45, -1e-45, 1.17549435e-38, 3.4028235e38, np.inf, -np.inf, np.nan, 1.1, 3.1415, 1e30, 1e-30], dtype=np.float32)
    ordinary = (rng.random(n, dtype=np.float32) * 4 - 2).astype(np.float32)
    a2 = ordinary.copy(); b2 = (rng.random(n, dtype=np.float32) + np.float32(0.5)).astype(np.float32)
    idx = rng.integers(0, n, 4096)
    a2[idx] = special[rng.integers(0, special.size, idx.size)]
    idx = rng.integers(0, n, 4096)
    b2[idx] = special[rng.integers(0, special.size, idx.size)]
    a2[::7] = 0.0                                   # out-of-the-money payoffs: zero numerators in every warp
    for x, y in ((a, b), (a2, b2)):
        X, Y = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, y)
        assert bits_equal(X.div(Y).getRealizationsFloat(), O.op_vv(O.DIV, x, y))
        assert bits_equal(X.vid(Y).getRealizationsFloat(), O.op_vv(O.VID, x, y))
        assert bits_equal(X.discount(Y, 0.5).getRealizationsFloat(), O.op_vvs(O.DISCOUNT, x, y, 0.5))
        # accumulator-side chains reach the _S / fused forms: (x*1) / y and y / (x*1)
        assert bits_equal(X.mult(1.0).add(0.0).div(Y).getRealizationsFloat(), O.op_vv(O.DIV, O.op_vs(O.ADD, O.op_vs(O.MULT, x, 1.0), 0.0), y))
        for s in (1.1, 3.0, 2.0 ** -60, 2.0 ** 60, 0.0, -0.0, float("inf"), float("nan"), 1e-45):
            assert bits_equal(X.div(s).getRealizationsFloat(), O.op_vs(O.DIV, x, s)), s
            assert bits_equal(X.vid(s).getRealizationsFloat(), O.op_vs(O.VID, x, s)), s


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 2047, 2048, 2049, 4097, 65536 + 3])
def test_ragged_sizes(fc, O, n):
    rng = np.random.RandomState(n + 1)
    x = rng.standard_normal(n); y = rng.standard_normal(n)
    X, Y = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, y)
    xf, yf = O.from_f64(x), O.from_f64(y)
    got = X.mult(Y).add(1.0).sub(X).getRealizationsFloat()
    want = O.op_vv(O.SUB, O.op_vs(O.ADD, O.op_vv(O.MULT, xf, yf), 1.0), xf)
    assert bits_equal(got, want)
    if n > 0:
        assert abs(X.mult(Y).getAverage() - O.average(O.op_vv(O.MULT, xf, yf))) <= 1e-12 + 1e-9 * abs(O.average(O.op_vv(O.MULT, xf, yf)))
    else:
        assert math.isnan(X.getAverage())


def test_reductions(fc, O, data):
    x, y, z = data
    for src in (x, z, x * 1000.0 + 5000.0):
        rv = fc.RandomVariableCuda(0.0, src)
        f = O.from_f64(src)
        for got, want in ((rv.getAverage(), O.average(f)), (rv.getVariance(), O.variance(f)),
                          (rv.getSampleVariance(), O.sample_variance(f)),
                          (rv.getStandardDeviation(), math.sqrt(O.variance(f))),
                          (rv.getStandardError(), math.sqrt(O.variance(f)) / math.sqrt(N))):
            assert abs(got - want) <= 1e-5 * abs(want), (got, want)
            assert abs(got - want) <= 1e-10 * abs(want) + 1e-300, ("tighter than required", got, want)
        assert rv.getMin() == O.minimum(f) and rv.getMax() == O.maximum(f)
    X, P = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, y / N)
    want = O.average(O.from_f64(x), O.from_f64(y / N))
    assert abs(X.getAverage(P) - want) <= 1e-10 * abs(want)
    want = O.variance(O.from_f64(x), O.from_f64(y / N))
    assert abs(X.getVariance(P) - want) <= 1e-9 * abs(want)
    # quantiles / histogram (RVF:472-602)
    for q in (0.0, 0.05, 0.5, 0.95, 1.0):
        assert X.getQuantile(q) == O.quantile(O.from_f64(x), q)
    assert abs(X.getQuantileExpectation(0.1, 0.9) - O.quantile_expectation(O.from_f64(x), 0.1, 0.9)) < 1e-12
    pts = np.linspace(0.1, 0.9, 9)
    assert np.allclose(X.getHistogram(pts), O.histogram(O.from_f64(x), pts), atol=0, rtol=0)


def test_order_statistics_on_the_device(fc, O):
    """getQuantile / getQuantileExpectation / getHistogram by radix select on the device must return exactly what
    Arrays.sort-based RVF:472-602 returns: ties, -0.0 < +0.0, NaN last, ragged sizes, boundary quantiles."""
    rng = np.random.default_rng(17)
    cases = [
        (rng.standard_normal(100_003)).astype(np.float32),
        np.round(rng.standard_normal(50_000) * 3).astype(np.float32),                 # heavy ties
        np.concatenate([rng.random(997).astype(np.float32), np.array([0.0, -0.0, -0.0, np.inf, -np.inf, np.nan, np.nan], dtype=np.float32)]),
        np.array([2.5], dtype=np.float32),
        np.full(4096, -1.25, dtype=np.float32),
    ]
    for x in cases:
        rng.shuffle(x)
        X = fc.RandomVariableCuda(0.0, x)
        for q in (0.0, 1e-6, 0.05, 0.25, 0.5, 0.75, 0.95, 0.999, 1.0):
            got, want = X.getQuantile(q), O.quantile(x, q)
            assert got == want or (got != got and want != want), (x.size, q, got, want)
        for q0, q1 in ((0.1, 0.9), (0.0, 1.0), (0.5, 0.5), (0.9, 0.1), (0.0, 0.0)):
            got, want = X.getQuantileExpectation(q0, q1), O.quantile_expectation(x, q0, q1)
            assert got == want or (got != got and want != want) or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (x.size, q0, q1, got, want)
        pts = np.array([-2.0, -0.5, 0.0, 0.3, 1.0, 4.0])
        assert np.array_equal(X.getHistogram(pts), O.histogram(x, pts)), x.size


def test_fused_chain_and_lazy_semantics(fc, O, data):
    x, y, z = data
    X, Y, Z = (fc.RandomVariableCuda(0.0, v) for v in (x, y, z))
    xf, yf, zf = O.from_f64(x), O.from_f64(y), O.from_f64(z)
    before = fc.stats()
    # Black-Scholes Euler step + payoff chain (MonteCarloBlackScholesModelTest.java:139-144), 7 ops, one getAverage
    s = X.add(0.005).addProduct(Z, 0.3).exp()
    payoff = s.sub(1.05).floor(0.0).div(1.1).mult(1.0)
    got = payoff.getAverage()
    after = fc.stats()
    so = O.op_v(O.EXP, O.op_vvs(O.ADDPRODUCT, O.op_vs(O.ADD, xf, 0.005), zf, 0.3))
    po = O.op_vs(O.MULT, O.op_vs(O.DIV, O.op_vs(O.FLOOR, O.op_vs(O.SUB, so, 1.05), 0.0), 1.1), 1.0)
    want = O.average(po)
    assert abs(got - want) <= 1e-6 * abs(want)
    assert after["n_tape_kernels"] - before["n_tape_kernels"] == 1, "the whole chain + reduction must be ONE kernel"
    # s is still referenced -> it was stored; payoff (the reduction target) stays pending and is recomputed on demand
    assert ulp_diff(s.getRealizationsFloat(), so) <= 1
    assert ulp_diff(payoff.getRealizationsFloat(), po) <= 1
    # dropping handles of pending nodes frees them without running anything
    k0 = fc.stats()["n_kernels"]
    t = X.mult(Y).add(Z).squared()
    del t
    assert fc.stats()["n_kernels"] == k0
    assert fc.stats()["pending_nodes"] == 0


def test_long_tape_register_pressure_and_cuts(fc, O, data):
    """A DAG with many live values (forces the shared-memory register file and HBM spills) and a chain long enough
    to be cut into several kernels must still be bit-exact."""
    x, y, z = data
    n = 20000
    xs = [np.float32(x[:n] + 0.01 * k) for k in range(24)]
    X = [fc.RandomVariableCuda(0.0, v) for v in xs]
    # 24 live intermediates, all consumed at the end
    inter = [X[k].mult(X[(k + 1) % 24]).add(float(k)) for k in range(24)]
    inter_o = [O.op_vs(O.ADD, O.op_vv(O.MULT, xs[k], xs[(k + 1) % 24]), float(k)) for k in range(24)]
    acc = inter[0]; acc_o = inter_o[0]
    for k in range(1, 24):
        acc = acc.addProduct(inter[k], inter[(k * 7) % 24]); acc_o = O.op_vvv(O.ADDPRODUCT, acc_o, inter_o[k], inter_o[(k * 7) % 24])
    del inter
    assert bits_equal(acc.getRealizationsFloat(), acc_o)
    # long chain: 3000 ops -> several tapes
    c = X[0]; co = xs[0]
    for k in range(1000):
        c = c.mult(1.0001).add(X[k % 24]).sub(0.5); co = O.op_vs(O.SUB, O.op_vv(O.ADD, O.op_vs(O.MULT, co, 1.0001), xs[k % 24]), 0.5)
    assert bits_equal(c.getRealizationsFloat(), co)


SCHED_DEFAULTS = {"ring_max": 16, "ring_min": 2, "target_ctas": 0, "horizon": 96, "pipeline": 1, "max_sets": 1, "grid_limit": 0, "max_regs": 8}


@pytest.mark.parametrize("opts", [
    {"grid_limit": 1}, {"grid_limit": 3, "max_sets": 2}, {"grid_limit": 2, "max_sets": 3, "ring_max": 2, "ring_min": 1},
    {"grid_limit": 5, "pipeline": 0}, {"ring_max": 3, "horizon": 4, "grid_limit": 7}, {"ring_max": 16, "target_ctas": 1, "grid_limit": 4},
    {"max_sets": 1, "grid_limit": 2}])
def test_interpreter_scheduling_variants_bit_exact(fc, O, data, opts):
    """Same tapes under different ring depths, slot-set counts (cross-chunk prefetch depth), pipelining on/off and grids
    small enough that every warp walks many chunks (prologue / T_LOADN / ragged last chunk): results must not move."""
    x, y, z = data
    n = 77_777
    xs = [np.float32(v[:n]) for v in (x, y, z)] + [np.float32(x[:n] * 0.5 + 0.25 * k) for k in range(9)]
    try:
        for k_, v_ in opts.items():
            fc.set_option(k_, v_)
        V = [fc.RandomVariableCuda(0.0, v) for v in xs]
        # many leaves through a small ring, a leaf reused far apart, two stored results and a fused weighted reduction
        acc = V[0].mult(V[1]); acc_o = O.op_vv(O.MULT, xs[0], xs[1])
        for k in range(2, 12):
            acc = acc.addProduct(V[k], 0.5).sub(V[(k * 5) % 12]); acc_o = O.op_vv(O.SUB, O.op_vvs(O.ADDPRODUCT, acc_o, xs[k], 0.5), xs[(k * 5) % 12])
        side = acc.mult(V[0]).floor(0.0); side_o = O.op_vs(O.FLOOR, O.op_vv(O.MULT, acc_o, xs[0]), 0.0)
        got_avg = acc.getAverage(V[2].abs())
        want_avg = O.average(acc_o, O.op_v(O.ABS, xs[2]))
        assert bits_equal(acc.getRealizationsFloat(), acc_o)
        assert bits_equal(side.getRealizationsFloat(), side_o)
        assert abs(got_avg - want_avg) <= 1e-6 * max(1.0, abs(want_avg))
        for red, ored in (("getAverage", O.average), ("getVariance", O.variance), ("getMin", O.minimum), ("getMax", O.maximum)):
            got = getattr(side, red)()
            want = ored(side_o)
            assert abs(got - want) <= 1e-5 * max(abs(want), 1e-30), (red, got, want)
    finally:
        for k_, v_ in SCHED_DEFAULTS.items():
            fc.set_option(k_, v_)


def test_unfused_mode_matches(fc, O, data):
    x, y, _ = data
    xf, yf = O.from_f64(x), O.from_f64(y)
    fc.set_option("fuse", 0)
    try:
        k0 = fc.stats()["n_tape_kernels"]
        X, Y = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, y)
        r = X.mult(Y).add(1.0).div(Y)
        assert fc.stats()["n_tape_kernels"] - k0 == 3        # the reference's execution model: one kernel per op
        assert bits_equal(r.getRealizationsFloat(), O.op_vv(O.DIV, O.op_vs(O.ADD, O.op_vv(O.MULT, xf, yf), 1.0), yf))
    finally:
        fc.set_option("fuse", 1)


def test_mt19937_stream_bit_exact(fc, O):
    from finmath_cuda.brownian_motion import mt19937_raw
    for mode in (O.SEED_LONG, O.SEED_INT):
        for seed in (31415, 5489, 1234):
            got = mt19937_raw(seed, 200_000, mode)
            assert np.array_equal(got, O.mt_u32(seed, 200_000, mode)), (mode, seed)
    # deep offsets exercise the jump-ahead polynomials (several chunk bits set)
    for skip in (1, 623, 624, 32768, 32768 * 5 + 17, 10_000_019):
        got = mt19937_raw(31415, 5000, O.SEED_LONG, skip)
        assert np.array_equal(got, O.mt_u32(31415, 5000, O.SEED_LONG, skip)), skip
    # mt19937ar known answers
    assert list(mt19937_raw(5489, 3, O.SEED_INT)) == [3499211612, 581869302, 3890346734]


@pytest.mark.parametrize("T,F,n,dt", [(100, 1, 5000, 1.0), (80, 1, 10000, 0.5), (40, 6, 4099, 0.5), (3, 2, 7, 0.1), (500, 1, 300, 0.01)])
def test_brownian_increments_bit_exact(fc, O, T, F, n, dt):
    td = fc.TimeDiscretization(0.0, T, dt)
    sqrt_dt = np.array([math.sqrt(td.getTimeStep(t)) for t in range(T)])
    for seed, mode in ((31415, O.SEED_LONG), (314151, O.SEED_INT)):
        bm = fc.BrownianMotionCuda(td, F, n, seed, seedMode=mode)
        want = O.brownian(seed, T, F, n, sqrt_dt, mode)
        for t in (0, T // 2, T - 1):
            for f in range(F):
                inc = bm.getBrownianIncrement(t, f)
                assert inc.getFiltrationTime() == td.getTime(t + 1)
                assert bits_equal(inc.getRealizationsFloat(), want[t * F + f]), (T, F, n, t, f)


def test_brownian_path_slices_concatenate(fc, O):
    """Multi-GPU sharding: a rank that owns paths [p0,p1) generates exactly the increments of those paths."""
    T, F, n = 20, 2, 10007
    td = fc.TimeDiscretization(0.0, T, 0.25)
    full = fc.BrownianMotionCuda(td, F, n, 53252)
    parts = [fc.BrownianMotionCuda(td, F, n, 53252, pathRange=fc.distributed.path_slice(n, r, 4)) for r in range(4)]
    for t in (0, 7, 19):
        for f in range(F):
            whole = full.getBrownianIncrement(t, f).getRealizationsFloat()
            cat = np.concatenate([p.getBrownianIncrement(t, f).getRealizationsFloat() for p in parts])
            assert bits_equal(whole, cat)


def test_brownian_moments(fc):
    """BrownianMotionTest.java:66-122: mean within 3*sqrt(dt)/sqrt(n), variance within 3*dt/sqrt(n) (1e6 paths, dt 0.1)."""
    n, dt = 1_000_000, 0.1
    bm = fc.BrownianMotionCuda(fc.TimeDiscretization(0.0, 10, dt), 1, n, 1234)
    inc = bm.getBrownianIncrement(0, 0)
    assert abs(inc.getAverage()) < 3.0 * math.sqrt(dt) / math.sqrt(n)
    assert abs(inc.getVariance() - dt) < 3.0 * dt / math.sqrt(n)


def _exact_normal_equations(basis_o, y):
    """The normal equations from EXACT products of the float-valued columns: what RandomVariableFromDoubleArray (finmath-lib's
    default CPU type) computes for the same inputs; deterministic x deterministic entries are the plain double product."""
    k, n = len(basis_o), len(y)
    cols = [np.full(n, float(np.float32(b)), dtype=np.float64) if np.isscalar(b) else np.asarray(b, dtype=np.float64) for b in basis_o]
    yd = np.asarray(y, dtype=np.float64)
    XtX = np.empty((k, k)); XtY = np.empty(k)
    for i in range(k):
        for j in range(k):
            XtX[i, j] = float(basis_o[i]) * float(basis_o[j]) if (np.isscalar(basis_o[i]) and np.isscalar(basis_o[j])) else math.fsum(cols[i] * cols[j]) / n
        XtY[i] = math.fsum(cols[i] * yd) / n
    return XtX, XtY


@pytest.fixture(params=[1, 0], ids=["float_products", "exact_products"])
def regression_mode(fc, request):
    """Option regression_float_products: 1 (default) sums the float products like RandomVariableFromFloatArray (the oracle);
    0 sums the exact products — RandomVariableFromDoubleArray's value for the same inputs: HBM-bound instead of compute-bound,
    but an ill-conditioned basis (k = 8 below) turns the ~1e-8 difference of the sums into > 1e-5 in the coefficients, outside
    the tolerance towards the float class. Opt-in."""
    fc.set_option("regression_float_products", request.param)
    yield request.param
    fc.set_option("regression_float_products", 1)


def _check_normal_equations(mode, XtX, XtY, basis_o, y, XtX_o, XtY_o, atol):
    if mode == 1:      # the float class's sums: as tight as the summation order allows
        assert np.allclose(XtX, XtX_o, rtol=1e-9, atol=atol) and np.allclose(XtY, XtY_o, rtol=1e-9, atol=atol)
    else:              # exact products: tight against their own definition, inside the north-star tolerance against the float class
        XtX_e, XtY_e = _exact_normal_equations(basis_o, y)
        assert np.allclose(XtX, XtX_e, rtol=1e-12, atol=1e-15) and np.allclose(XtY, XtY_e, rtol=1e-12, atol=1e-15)
        scale = max(1.0, float(np.max(np.abs(XtX_o))))
        assert np.allclose(XtX, XtX_o, rtol=1e-6, atol=1e-7 * scale) and np.allclose(XtY, XtY_o, rtol=1e-6, atol=1e-7 * scale)


@pytest.mark.parametrize("k", [3, 6, 8])
def test_regression_normal_equations(fc, O, data, k, regression_mode):
    x, y, z = data
    xf, yf, zf = O.from_f64(x), O.from_f64(y), O.from_f64(z)
    X, Y, Z = (fc.RandomVariableCuda(0.0, v) for v in (x, y, z))
    one = fc.RandomVariableCuda(1.0)
    basis = [one, X, X.squared(), Y, Y.squared(), X.mult(Y), X.pow(3.0), Y.exp()][:k]
    basis_o = [1.0, xf, O.op_v(O.SQUARED, xf), yf, O.op_v(O.SQUARED, yf), O.op_vv(O.MULT, xf, yf), O.op_vs(O.POW, xf, 3.0), O.op_v(O.EXP, yf)][:k]
    from finmath_cuda.conditional_expectation import normal_equations
    XtX, XtY = normal_equations(basis, Z)
    XtX_o, XtY_o = O.regression_normal_eq(basis_o, zf)
    _check_normal_equations(regression_mode, XtX, XtY, basis_o, zf, XtX_o, XtY_o, 1e-14)
    est = fc.MonteCarloConditionalExpectationRegression(basis)
    c = est.getLinearRegressionParameters(Z)
    c_o = np.linalg.lstsq(XtX_o, XtY_o, rcond=None)[0]
    if regression_mode == 1:
        assert np.allclose(c, c_o, rtol=1e-5, atol=1e-8)
    else:
        c_e = np.linalg.lstsq(*_exact_normal_equations(basis_o, zf), rcond=None)[0]
        assert np.allclose(c, c_e, rtol=1e-5, atol=1e-8)
        assert np.allclose(c, c_o, rtol=1e-3, atol=1e-6)
    ce = Z.getConditionalExpectation(est)
    want = O.op_vs(O.MULT, np.full(N, 1.0, dtype=np.float32), c_o[0]) if False else None
    # estimate = basis[0]*c0 + sum c_i*basis[i] (float arithmetic); compare against the oracle evaluation of the same chain
    acc = np.full(N, np.float32(c_o[0]), dtype=np.float32) if np.isscalar(basis_o[0]) else O.op_vs(O.MULT, basis_o[0], c_o[0])
    for i in range(1, k):
        acc = O.op_vvs(O.ADDPRODUCT, acc, basis_o[i], c_o[i])
    got = ce.getRealizationsFloat()
    assert np.allclose(got, acc, rtol=1e-4, atol=1e-5)


def test_size_mismatch_and_bad_handles(fc):
    a = fc.RandomVariableCuda(0.0, np.arange(10.0))
    b = fc.RandomVariableCuda(0.0, np.arange(11.0))
    with pytest.raises(IndexError):
        a.add(b)
    with pytest.raises(IndexError):
        a.get(10)


def test_pool_recycles_and_purges(fc):
    """BrownianMotionMemoryTest.java:40-80 in spirit: many motions of growing size do not exhaust memory, buffers are recycled."""
    fc.RandomVariableCuda.purge()
    s0 = fc.stats()
    for i in range(20):
        bm = fc.BrownianMotionCuda(fc.TimeDiscretization(0.0, 10, 0.1), 1, 100_000 + 10_000 * i, 53252)
        inc = bm.getBrownianIncrement(3, 0)
        assert abs(inc.getVariance() - 0.1) < 5.0 * 0.1 / math.sqrt(100_000)
        del bm, inc
    s1 = fc.stats()
    assert s1["bytes_in_use"] <= s0["bytes_in_use"] + (1 << 20)
    x = [fc.RandomVariableCuda(0.0, np.zeros(50_000)) for _ in range(8)]
    del x
    y = [fc.RandomVariableCuda(0.0, np.zeros(50_000)) for _ in range(8)]
    s2 = fc.stats()
    assert s2["n_alloc_reused"] > s1["n_alloc_reused"]
    del y
    fc.RandomVariableCuda.purge()
    assert fc.stats()["bytes_cached"] <= s2["bytes_cached"]


def test_host_threads_share_the_runtime(fc, O):
    """RandomVariableCuda objects are immutable and usable from any thread (RandomVariableCuda.java:64-65; the library's
    calibration values its products from a thread pool). Four threads record, reduce and read back concurrently; every
    thread's results must be the single-threaded ones. A waiting reduction does not hold the runtime lock."""
    import threading
    n = 50_000
    rng = np.random.default_rng(99)
    inputs = [(rng.uniform(0.5, 2.0, n), rng.uniform(-1.0, 1.0, n)) for _ in range(4)]

    def work(x, y, rounds):
        out = []
        X, Y = fc.RandomVariableCuda(0.0, x), fc.RandomVariableCuda(0.0, y)
        for r in range(rounds):
            v = X.mult(1.0 + 0.125 * r).add(Y).discount(X, 0.5).floor(0.0)
            out.append((v.getAverage(), v.squared().getMax(), X.getAverage(), float(v.getRealizationsFloat()[r])))
        return out

    want = [work(x, y, 12) for x, y in inputs]
    got = [None] * 4
    errors = []

    def run(k):
        try:
            got[k] = work(*inputs[k], 12)
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errors, errors
    assert got == want
    # and against the oracle for one of them
    x, y = inputs[0]
    xf, yf = O.from_f64(x), O.from_f64(y)
    v = O.op_vs(O.FLOOR, O.op_vvs(O.DISCOUNT, O.op_vv(O.ADD, O.op_vs(O.MULT, xf, 1.0), yf), xf, 0.5), 0.0)
    assert abs(want[0][0][0] - O.average(v)) <= 1e-9 * abs(O.average(v))


def _random_program(rs, rv, leaves, nops=None, transcendental=False):
    """One random expression DAG over `leaves` (product objects or _OracleRV): the STRUCTURE comes from rs, every scalar
    from rv. Covers every fused form of the code generator, kept intermediates (stored), and six reductions."""
    vals = list(leaves)
    kept = []
    if nops is None: nops = int(rs.integers(25, 70))
    for _ in range(nops):
        a = vals[int(rs.integers(max(0, len(vals) - 8), len(vals)))]
        b = vals[int(rs.integers(0, len(vals)))]
        c = vals[int(rs.integers(0, len(vals)))]
        s1 = float(rv.uniform(0.25, 1.75)); s2 = float(rv.uniform(-0.5, 0.5))
        if rs.integers(0, 10) == 0: s1 = 1.0                   # the ACCRUE / DISCOUNT fusion looks for x * p + 1
        k = int(rs.integers(0, 28 if transcendental else 24))
        if k == 0: r = a.add(s2)
        elif k == 1: r = a.sub(s2).mult(s1)
        elif k == 2: r = a.mult(s1).add(s2)
        elif k == 3: r = a.add(b)
        elif k == 4: r = a.sub(b)
        elif k == 5: r = a.mult(b)
        elif k == 6: r = a.div(b.abs().add(0.5))
        elif k == 7: r = a.accrue(b, s1)
        elif k == 8: r = a.discount(b.abs(), s1)
        elif k == 9: r = a.addProduct(b, s2)
        elif k == 10: r = a.addProduct(b, c)
        elif k == 11: r = a.sub(s2).choose(b, c)
        elif k == 12: r = a.floor(s2).cap(s1)
        elif k == 13: r = a.squared().add(1.0).vid(s1)
        elif k == 14: r = b.mult(s1).add(1.0).mult(a)          # accrue written out
        elif k == 15: r = b.abs().mult(s1).add(1.0).vid(a)     # discount written out: a / (1 + |b| s1)
        elif k == 16: r = a.bus(s2).abs().sqrt()
        elif k == 17: r = a.mult(s1).add(s2).vid(0.5).mult(s1) # RATIO
        elif k == 18: r = a.mult(s1).add(s2).mult(0.75)        # MULADDMUL
        elif k == 19: r = a.add(b.sub(s2).mult(s1)).discount(b, s1)   # ADDAFFDISC
        elif k == 20: r = a.mult(a).add(a)
        elif k == 21: r = a.add(b.sub(s2).mult(s1))            # ADDAFF
        elif k == 22: r = a.cap(4.0).floor(-4.0)
        elif k == 23: r = a.sub(b).abs().sqrt().sub(s1).choose(a, c)
        elif k == 24: r = a.cap(2.0).exp()
        elif k == 25: r = a.abs().add(0.1).log()
        elif k == 26: r = a.abs().add(0.25).invert()
        else: r = a.abs().pow(float(rs.choice([2.0, 0.5, 3.0, 1.0])))
        vals.append(r)
        if rs.integers(0, 4) == 0: kept.append(r)              # a handle the caller keeps: must be stored
    out = [v.getRealizationsFloat().copy() for v in kept[-4:]] + [vals[-1].getRealizationsFloat().copy()]
    w = vals[0].abs()
    red = (vals[-2].getAverage(), vals[-3].getVariance(), vals[-4].getMax(), vals[-5].getMin() if len(vals) >= 5 else 0.0,
           vals[-3].getAverage(w), vals[-2].getVariance(w))
    return out, red


class _OracleRV:
    """The oracle behind the RandomVariable method names _random_program uses (stochastic values only)."""
    def __init__(self, O, v): self.O, self.v = O, v
    def _s(self, op, s): return _OracleRV(self.O, self.O.op_vs(op, self.v, s))
    def _v(self, op, o): return _OracleRV(self.O, self.O.op_vv(op, self.v, o.v))
    def _u(self, op): return _OracleRV(self.O, self.O.op_v(op, self.v))
    def _b(self, op, o): return self._v(op, o) if isinstance(o, _OracleRV) else self._s(op, o)
    def add(self, o): return self._b(self.O.ADD, o)
    def sub(self, o): return self._b(self.O.SUB, o)
    def bus(self, o): return self._b(self.O.BUS, o)
    def mult(self, o): return self._b(self.O.MULT, o)
    def div(self, o): return self._b(self.O.DIV, o)
    def vid(self, o): return self._b(self.O.VID, o)
    def floor(self, o): return self._b(self.O.FLOOR, o)
    def cap(self, o): return self._b(self.O.CAP, o)
    def pow(self, e): return self._s(self.O.POW, e)
    def abs(self): return self._u(self.O.ABS)
    def sqrt(self): return self._u(self.O.SQRT)
    def squared(self): return self._u(self.O.SQUARED)
    def exp(self): return self._u(self.O.EXP)
    def log(self): return self._u(self.O.LOG)
    def invert(self): return self._u(self.O.INVERT)
    def accrue(self, r, p): return _OracleRV(self.O, self.O.op_vvs(self.O.ACCRUE, self.v, r.v, p))
    def discount(self, r, p): return _OracleRV(self.O, self.O.op_vvs(self.O.DISCOUNT, self.v, r.v, p))
    def addProduct(self, a, b):
        if isinstance(b, _OracleRV): return _OracleRV(self.O, self.O.op_vvv(self.O.ADDPRODUCT, self.v, a.v, b.v))
        return _OracleRV(self.O, self.O.op_vvs(self.O.ADDPRODUCT, self.v, a.v, b))
    def choose(self, a, b): return _OracleRV(self.O, self.O.op_vvv(self.O.CHOOSE, self.v, a.v, b.v))
    def getRealizationsFloat(self): return self.v
    def getAverage(self, w=None): return self.O.average(self.v, None if w is None else w.v)
    def getVariance(self, w=None): return self.O.variance(self.v, None if w is None else w.v)
    def getMax(self): return self.O.maximum(self.v)
    def getMin(self): return self.O.minimum(self.v)


def _same_reduction(g, w):
    # (a non-finite element: the reference's Kahan loop turns inf into NaN (RVF:322-330) unless it comes last; getAverage mirrors that,
    # the other reductions may keep the infinity)
    return (g != g and w != w) or (w != w and abs(g) == float("inf")) or g == w or abs(g - w) <= 1e-9 * abs(w) + 1e-300


def test_average_of_a_vector_with_infinities_follows_the_kahan_loop(fc, O):
    """RVF:322-330: an infinity anywhere but at the last index turns the compensated sum into NaN; at the last index it survives."""
    base = np.linspace(-1.0, 2.0, 3000).astype(np.float32)
    for pos, val in ((17, np.inf), (2999, np.inf), (2999, -np.inf), (0, -np.inf), (1500, np.inf)):
        x = base.copy(); x[pos] = val
        got, want = fc.RandomVariableCuda(0.0, x.astype(np.float64)).getAverage(), O.average(x)
        assert (got != got and want != want) or got == want, (pos, val, got, want)
        got2 = fc.RandomVariableCuda(0.0, x.astype(np.float64)).mult(1.0).add(0.0).getAverage()       # through a fused chain
        assert (got2 != got2 and want != want) or got2 == want, (pos, val, got2, want)
    x = base.copy(); x[5] = np.inf; x[2999] = np.inf
    assert math.isnan(fc.RandomVariableCuda(0.0, x.astype(np.float64)).getAverage()) and math.isnan(O.average(x))


def test_fuzzed_programs_match_the_oracle(fc, O):
    """Random expression DAGs against the oracle: sizes around the 512-path chunk, 10 to 1200 operations (long ones cross
    the automatic flush, the tape limits and the register file), 1 to 8 leaves, the scheduler knobs of the variants test,
    every structure with three sets of values (the second and third are replays from the tape cache). On the GPU the
    double-then-round transcendentals are left out (1 ulp there would be amplified by what follows); on the emulator
    (tests/test_codegen_emulator.py) they are in."""
    import os
    emulated = bool(os.environ.get("FMC_TEST_TAPE_EMULATOR"))
    knobs = [{}, {"ring_max": 2, "ring_min": 1}, {"grid_limit": 1}, {"grid_limit": 2, "max_sets": 2}, {"pipeline": 0}, {"horizon": 4}, {"target_ctas": 1},
             {"max_regs": 4}, {"max_regs": 16}]
    try:
        for opts in knobs:
            for k_, v_ in SCHED_DEFAULTS.items(): fc.set_option(k_, v_)
            for k_, v_ in opts.items(): fc.set_option(k_, v_)
            for seed in range(int(os.environ.get("FMC_FUZZ_SEEDS", "24"))):      # more seeds for an occasional long run
                rs0 = np.random.default_rng(1000 * len(opts) + seed)
                n = int(rs0.choice([1, 5, 100, 511, 512, 513, 700, 1500, 3000, 5000]))
                nops = int(rs0.choice([10, 40, 120, 400, 1200]))
                nleaf = int(rs0.integers(1, 9))
                for vseed in range(3):
                    xs = [np.random.default_rng(seed * 7 + i + 1000 * vseed).uniform(0.2, 2.0, n) for i in range(nleaf)]
                    got = _random_program(np.random.default_rng(seed), np.random.default_rng(seed + 99 + vseed),
                                          [fc.RandomVariableCuda(0.0, x) for x in xs], nops, emulated)
                    want = _random_program(np.random.default_rng(seed), np.random.default_rng(seed + 99 + vseed),
                                           [_OracleRV(O, O.from_f64(x)) for x in xs], nops, emulated)
                    assert all(bits_equal(g, w) for g, w in zip(got[0], want[0])), (opts, seed, vseed, n, nops, nleaf)
                    assert all(_same_reduction(g, w) for g, w in zip(got[1], want[1])), (opts, seed, vseed, n, nops, nleaf, got[1], want[1])
    finally:
        for k_, v_ in SCHED_DEFAULTS.items(): fc.set_option(k_, v_)


@pytest.mark.parametrize("n", [5, 4099, 70001])
def test_batched_averages_equal_single_reductions(fc, O, n):
    """A caller that builds many result vectors first and averages them afterwards (finmath-lib's calibration objective:
    all product values, then value.getAverage() in a second loop): the vectors one flush materialised together are summed
    in ONE launch at the first getAverage() and the other sums are handed out by the calls that follow. The sums must be
    bit-identical to the single-vector reduction, survive released / re-used handles, and never serve a consumed vector."""
    rng = np.random.default_rng(n)
    base = [rng.standard_normal(n) for _ in range(3)]

    def products(leaves, k):
        return [leaves[j % 3].mult(0.25 + 0.125 * j).add(leaves[(j + 1) % 3]).floor(-0.5).div(1.0 + 0.5 * j) for j in range(k)]

    try:
        results = {}
        for batch in (0, 1):
            fc.set_option("batch_reduce", batch)
            leaves = [fc.RandomVariableCuda(0.0, b) for b in base]
            k0 = fc.stats()["n_kernels"]
            vs = products(leaves, 12)
            fc.flush()
            k1 = fc.stats()["n_kernels"]
            got = [v.getAverage() for v in vs[:6]]
            del vs[7]                                   # released before its turn: the slot may be re-used by the next vector
            w = leaves[0].mult(3.0)                     # recorded between two averages: not part of any batch
            got.append(w.getAverage())
            got += [v.getAverage() for v in vs[6:]]
            got.append(vs[0].getAverage())              # asked twice
            c = vs[1].add(1.0)                          # a consumer of a batched vector: its own average is a new reduction
            got.append(c.getAverage())
            got.append(vs[1].getAverage())
            k2 = fc.stats()["n_kernels"]
            results[batch] = got
            if batch:
                assert k2 - k1 <= 6, "12 averages after one flush must not cost 12 launches"
            assert k1 > k0
        assert all(a == b for a, b in zip(results[0], results[1])), (results[0], results[1])
        # against the oracle
        fl = [O.from_f64(b) for b in base]
        for j in range(6):
            po = O.op_vs(O.DIV, O.op_vs(O.FLOOR, O.op_vv(O.ADD, O.op_vs(O.MULT, fl[j % 3], 0.25 + 0.125 * j), fl[(j + 1) % 3]), -0.5), 1.0 + 0.5 * j)
            want = O.average(po)
            assert abs(results[1][j] - want) <= 1e-10 * abs(want) + 1e-300, (j, results[1][j], want)
        # a series of averages whose tail was never flushed: the first is a fused chain -> reduce, results stay the same
        fc.set_option("batch_reduce", 1)
        leaves = [fc.RandomVariableCuda(0.0, b) for b in base]
        vs = products(leaves, 5)
        lazy = [v.getAverage() for v in vs]
        assert all(abs(a - b) <= 1e-12 * abs(b) for a, b in zip(lazy, results[1][:5]))
    finally:
        fc.set_option("batch_reduce", 1)


def test_tape_cache_replays_are_bit_exact(fc):
    """a cone whose structure was lowered before is replayed from the tape cache with the new buffers and
    immediates patched in. Replays must be indistinguishable from fresh code generation: same structure, three sets of
    scalars and inputs, cache off against cache on (miss, hit, hit)."""
    import ctypes
    from finmath_cuda import _capi as capi

    def counter(key):
        v = ctypes.c_double(); capi.check(capi.load().fmc_get_option(key.encode(), ctypes.byref(v))); return int(v.value)

    n = 3000
    try:
        for prog in range(12):
            results = {}
            for cache in (0, 1):
                fc.set_option("tape_cache", cache)
                for vset in (0, 1, 2, 0):
                    rv = np.random.default_rng(1000 * prog + vset)
                    leaves = [fc.RandomVariableCuda(0.0, rv.uniform(0.2, 2.0, n)) for _ in range(4)]
                    got = _random_program(np.random.default_rng(prog), rv, leaves)
                    if cache == 0: results[vset] = got
                    else:
                        want = results[vset]
                        assert all(bits_equal(g, w) for g, w in zip(got[0], want[0])), (prog, vset)
                        assert all(a == b or (a != a and b != b) for a, b in zip(got[1], want[1])), (prog, vset, got[1], want[1])
        assert counter("tape_cache_hits") > 100
    finally:
        fc.set_option("tape_cache", 1)


def test_uploads_on_the_copy_stream_do_not_overtake_queued_kernels(fc, O):
    """Host uploads run on their own stream, concurrently with kernels already queued. A block freed while a queued
    kernel may still read it must not be overwritten by the next upload: upload a, queue a long chain on it without
    waiting, drop a, upload b (which would recycle a's block), repeat; every chain result must come from ITS input."""
    from finmath_cuda import _capi as capi
    n = 6_000_000
    rng = np.random.default_rng(5)
    inputs = [rng.uniform(0.5, 1.5, n) for _ in range(3)]
    want = []
    for a in inputs:
        v = O.from_f64(a)
        for _ in range(6):
            v = O.op_v(O.SQRT, O.op_vs(O.ADD, O.op_vs(O.MULT, v, 1.25), 0.5))
        want.append(v)
    for round_ in range(3):
        results = []
        for a in inputs:
            x = fc.RandomVariableCuda(0.0, a)
            y = x
            for _ in range(6):
                y = y.mult(1.25).add(0.5).sqrt()
            z = y.add(0.0)                       # force y to be stored by a launch that is only QUEUED here
            capi.check(capi.load().fmc_flush())
            del x, y                             # frees the input (and intermediate) blocks while the kernel may still run
            results.append(z)
        for z, w in zip(results, want):
            assert bits_equal(z.getRealizationsFloat(), O.op_vs(O.ADD, w, 0.0)), round_


def test_fuzzed_statistics_regression_and_brownian(fc, O):
    """Randomised sizes and parameters for the kernels next to the interpreter: radix-select order statistics (ties,
    signed zeros, infinities), histogram, the fused regression normal equations (k = 1..12 with deterministic columns)
    and Brownian path slices with random (T, F, n, p0, p1, seed)."""
    from finmath_cuda.conditional_expectation import normal_equations
    rng = np.random.default_rng(2024)
    for case in range(40):
        n = int(rng.choice([1, 2, 31, 255, 256, 257, 1023, 4097, 65537, 300_001]))
        kind = case % 4
        if kind == 0: x = rng.standard_normal(n)
        elif kind == 1: x = np.round(rng.standard_normal(n) * 2.0) / 2.0              # ties
        elif kind == 2: x = np.where(rng.random(n) < 0.3, 0.0, rng.standard_normal(n)) * np.where(rng.random(n) < 0.5, -1.0, 1.0)   # +-0
        else: x = np.concatenate([rng.standard_normal(max(n - 2, 0)), [np.inf, -np.inf][: min(n, 2)]])[:n]
        x = x.astype(np.float32)
        X = fc.RandomVariableCuda(0.0, x)
        for q in list(rng.random(3)) + [0.0, 1.0]:
            got, want = X.getQuantile(float(q)), O.quantile(x, float(q))
            assert got == want, (case, n, q, got, want)
        q0, q1 = sorted(rng.random(2))
        got, want = X.getQuantileExpectation(float(q0), float(q1)), O.quantile_expectation(x, float(q0), float(q1))
        assert (got != got and want != want) or got == want or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (case, n, q0, q1, got, want)
        pts = np.sort(rng.standard_normal(int(rng.integers(1, 9))))
        assert np.array_equal(X.getHistogram(pts), O.histogram(x, pts)), (case, n)
    for case in range(12):
        n = int(rng.choice([7, 513, 5000, 100_003]))
        k = int(rng.integers(1, 13))
        cols = [rng.uniform(-1.0, 1.0, n).astype(np.float32) for _ in range(k)]
        det = [bool(rng.random() < 0.2) for _ in range(k)]
        y = rng.uniform(-1.0, 1.0, n).astype(np.float32)
        basis = [fc.RandomVariableCuda(0.0, float(c[0])) if d else fc.RandomVariableCuda(0.0, c.astype(np.float64)) for c, d in zip(cols, det)]
        basis_o = [float(np.float64(c[0])) if d else c for c, d in zip(cols, det)]
        XtX_o, XtY_o = O.regression_normal_eq(basis_o, y)
        for mode in (0, 1):
            fc.set_option("regression_float_products", mode)
            try:
                XtX, XtY = normal_equations(basis, fc.RandomVariableCuda(0.0, y.astype(np.float64)))
            finally:
                fc.set_option("regression_float_products", 1)
            _check_normal_equations(mode, XtX, XtY, basis_o, y, XtX_o, XtY_o, 1e-13)
    for case in range(10):
        T, F = int(rng.integers(1, 30)), int(rng.integers(1, 5))
        n = int(rng.choice([1, 33, 1000, 20_011]))
        p0 = int(rng.integers(0, n)); p1 = int(rng.integers(p0, n + 1))
        seed, mode = int(rng.integers(1, 2**31 - 1)), int(rng.integers(0, 2))
        dt = float(rng.choice([0.1, 0.25, 1.0]))
        td = fc.TimeDiscretization(0.0, T, dt)
        sqrt_dt = np.array([math.sqrt(td.getTimeStep(t)) for t in range(T)])
        bm = fc.BrownianMotionCuda(td, F, n, seed, seedMode=mode, pathRange=(p0, p1))
        want = O.brownian(seed, T, F, n, sqrt_dt, mode, p0, p1)
        for t in {0, T // 2, T - 1}:
            for f in range(F):
                got = bm.getBrownianIncrement(t, f).getRealizationsFloat() if p1 > p0 else np.empty(0, dtype=np.float32)
                assert bits_equal(got, want[t * F + f]), (case, T, F, n, p0, p1, seed, mode, t, f)


def test_pinned_asynchronous_upload(fc, O):
    """fmc_host_alloc + fmc_vec_from_f64_pinned: DMA of the doubles on the copy stream and (float) cast on the device must
    give exactly the vector the synchronous upload gives (Java's (float) cast, RVC:768-774), for sizes around the 1 Mi
    element chunk, while kernels are queued; buffers that are not pinned are refused."""
    import ctypes as C
    from finmath_cuda import _capi as capi
    L = capi.load()
    rng = np.random.default_rng(11)
    for n in (1, 7, 4097, (1 << 20) - 1, (1 << 20) + 3, 3 * (1 << 20) + 17):
        p = C.c_void_p()
        capi.check(L.fmc_host_alloc(8 * n, C.byref(p)))
        host = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))
        host[:] = rng.standard_normal(n) * np.where(rng.random(n) < 0.01, 1e-40, 1.0)      # some values denormal as floats
        busy = fc.RandomVariableCuda(0.0, rng.random(2_000_000))
        for _ in range(8): busy = busy.mult(1.0001).add(0.5).sqrt()
        keep = busy.add(0.0); capi.check(L.fmc_flush())                                     # kernels in flight during the upload
        h = C.c_uint64()
        capi.check(L.fmc_vec_from_f64_pinned(p, n, C.byref(h)))
        out = np.empty(n, dtype=np.float32)
        capi.check(L.fmc_vec_to_f32(h.value, out.ctypes.data, n))
        assert bits_equal(out, O.from_f64(host)), n
        capi.check(L.fmc_vec_release(h.value))
        del keep, busy
        capi.check(L.fmc_host_free(p))
    x = rng.random(100)
    h = C.c_uint64()
    assert L.fmc_vec_from_f64_pinned(x.ctypes.data, 100, C.byref(h)) == capi.FMC_ERR_INVALID
