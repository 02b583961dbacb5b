"""Config 4 of BASELINE.json: delta / vega by backward-mode AAD with RandomVariableDifferentiableAAD wrapping
RandomVariableCuda (README.md:50-52 of the reference: the AAD wrapper composes with the GPU type and has the higher
type priority). Every primal AND adjoint operation goes through RandomVariableCuda's own methods, i.e. through the C ABI
and the fused interpreter; gradients are checked against finite differences of the same GPU computation, against
numpy, and against the Black-Scholes closed forms."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_type_priority_aad_over_gpu_over_cpu(fc):
    x = fc.RandomVariableCuda(0.0, np.linspace(0.5, 1.5, 1000))
    ax = fc.RandomVariableDifferentiableAAD(x)
    assert ax.getTypePriority() > x.getTypePriority() == 20
    # a plain GPU vector hands the operation over to the differentiable operand (mirror methods, RVF:962-1178)
    for name in ("add", "sub", "mult", "div", "cap", "floor"):
        r = getattr(x, name)(ax)
        assert isinstance(r, fc.RandomVariableDifferentiableAAD), name
        want = getattr(x, name)(x).getRealizationsFloat()
        assert np.array_equal(r.getRealizations().astype(np.float32), want), name
    assert isinstance(x.accrue(ax, 0.5), fc.RandomVariableDifferentiableAAD)
    assert isinstance(x.discount(ax, 0.5), fc.RandomVariableDifferentiableAAD)
    assert isinstance(x.addProduct(ax, 2.0), fc.RandomVariableDifferentiableAAD)


def test_gradient_of_an_op_chain_matches_finite_differences(fc):
    n = 50_000
    rng = np.random.default_rng(5)
    x0 = (rng.random(n) + 0.5); y0 = (rng.random(n) + 0.5)

    def f_np(x, y):
        u = x * y + np.exp(0.3 * x) - np.log(y) / (1.0 + 0.5 * x)
        v = np.sqrt(u * u + 1.0) * (1.0 + y * 0.25)             # accrue(y, 0.25)
        w = np.maximum(v - 2.0, 0.0) + np.minimum(x, 1.0) + 1.0 / y
        return np.where(x - 1.0 >= 0, w * x, w + y * y)         # choose

    def f_rv(X, Y):
        u = X.mult(Y).add(X.mult(0.3).exp()).sub(Y.log().div(X.mult(0.5).add(1.0)))
        v = u.squared().add(1.0).sqrt().accrue(Y, 0.25)
        w = v.sub(2.0).floor(0.0).add(X.cap(1.0)).add(Y.invert())
        return X.sub(1.0).choose(w.mult(X), w.addProduct(Y, Y))

    X = fc.RandomVariableDifferentiableAAD(fc.RandomVariableCuda(0.0, x0))
    Y = fc.RandomVariableDifferentiableAAD(fc.RandomVariableCuda(0.0, y0))
    F = f_rv(X, Y)
    x32, y32 = x0.astype(np.float32).astype(np.float64), y0.astype(np.float32).astype(np.float64)
    assert np.allclose(F.getRealizations(), f_np(x32, y32), rtol=2e-5, atol=2e-5)
    g = F.getGradient()
    h = 1e-6
    dfdx = (f_np(x32 + h, y32) - f_np(x32 - h, y32)) / (2 * h)
    dfdy = (f_np(x32, y32 + h) - f_np(x32, y32 - h)) / (2 * h)
    # away from the kinks of floor / cap / choose
    smooth = (np.abs(x32 - 1.0) > 1e-3) & (np.abs(np.sqrt((x32 * y32 + np.exp(0.3 * x32) - np.log(y32) / (1.0 + 0.5 * x32)) ** 2 + 1.0) * (1.0 + y32 * 0.25) - 2.0) > 1e-3)
    gx, gy = g[X.getID()].getRealizations(), g[Y.getID()].getRealizations()
    assert np.allclose(gx[smooth], dfdx[smooth], rtol=2e-3, atol=2e-3)
    assert np.allclose(gy[smooth], dfdy[smooth], rtol=2e-3, atol=2e-3)


def _norm_cdf(x): return 0.5 * (1.0 + math.erf(x / math.sqrt(2.0)))


def test_black_scholes_delta_and_vega_by_aad_on_the_gpu(fc):
    """MonteCarloBlackScholesModelTest.java:62-76 constants; Euler scheme in log-coordinates over BrownianMotionCuda."""
    n, steps, T = 200_000, 20, 2.0
    S0, r, sigma, K = 1.0, 0.05, 0.30, 1.05
    dt = T / steps
    td = fc.TimeDiscretization(0.0, steps, dt)
    bm = fc.BrownianMotionCuda(td, 1, n, 31415)
    fac = fc.RandomVariableDifferentiableAADFactory(fc.RandomVariableCudaFactory())

    def price(s0, sig):
        x = s0.log()
        drift = sig.squared().mult(-0.5).add(r).mult(dt)
        for t in range(steps):
            x = x.add(drift).add(sig.mult(bm.getBrownianIncrement(t, 0)))
        payoff = x.exp().sub(K).floor(0.0).mult(math.exp(-r * T))
        return payoff.average()

    s0, sig = fac.createRandomVariable(0.0, S0), fac.createRandomVariable(0.0, sigma)
    k0 = fc.stats()["n_tape_kernels"]
    V = price(s0, sig)
    grad = V.getGradient()
    delta, vega = grad[s0.getID()].getAverage(), grad[sig.getID()].getAverage()
    kernels = fc.stats()["n_tape_kernels"] - k0
    d1 = (math.log(S0 / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * math.sqrt(T))
    delta_bs, vega_bs = _norm_cdf(d1), S0 * math.sqrt(T) * math.exp(-0.5 * d1 * d1) / math.sqrt(2 * math.pi)
    value_bs = S0 * _norm_cdf(d1) - K * math.exp(-r * T) * _norm_cdf(d1 - sigma * math.sqrt(T))
    assert abs(V.doubleValue() - value_bs) < 0.005                     # MonteCarloBlackScholesModelTest.java:156
    assert abs(delta - delta_bs) < 0.01 and abs(vega - vega_bs) < 0.02
    # against bump-and-revalue of the SAME GPU computation (same Brownian paths): pathwise derivative == finite difference
    plain = fc.RandomVariableCudaFactory()

    def price_plain(s0v, sigv):
        class W:                                                       # plain GPU variables through the same code
            pass
        return price(fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, s0v)),
                     fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, sigv))).doubleValue()
    h = 1e-3
    assert abs(delta - (price_plain(S0 + h, sigma) - price_plain(S0 - h, sigma)) / (2 * h)) < 2e-3
    assert abs(vega - (price_plain(S0, sigma + h) - price_plain(S0, sigma - h)) / (2 * h)) < 3e-3
    assert kernels < 40, f"primal + adjoint sweep should fuse into a few launches, used {kernels}"


def test_lmm_swaption_vega_by_aad_on_the_gpu(fc):
    """BASELINE config 4, LMM half: a small LIBOR market model (spot measure, normal state space, one factor, as
    drivers/workloads.hpp) simulated on RandomVariableDifferentiableAAD over the GPU type; the derivative of a swaption
    value with respect to the (flat) volatility parameter by one reverse sweep must equal bump-and-revalue on the same
    Brownian paths, and the sweep must stay a handful of fused launches."""
    n, NP, delta = 60_000, 8, 0.5
    L0, strike, sigma0 = 0.02, 0.02, 0.006
    td = fc.TimeDiscretization(0.0, NP, delta)
    bm = fc.BrownianMotionCuda(td, 1, n, 31415)
    plain = fc.RandomVariableCudaFactory()

    def swaption_value(sig):
        libor = [plain.createRandomVariable(0.0, L0) for _ in range(NP)]
        numeraire = [plain.createRandomVariable(0.0, 1.0)]
        exercise = NP // 2
        at_exercise = None
        for t in range(NP):
            if t == exercise: at_exercise = list(libor)
            numeraire.append(numeraire[-1].accrue(libor[t], delta))                    # N(T_{t+1}) = N(T_t) (1 + L_t(T_t) delta)
            dW = bm.getBrownianIncrement(t, 0)
            factor_sum = None
            new = list(libor)
            for i in range(t + 1, NP):
                transform = sig.mult(libor[i].mult(delta).add(1.0).invert().mult(delta))   # sigma * delta / (1 + delta L_i)
                factor_sum = transform if factor_sum is None else factor_sum.add(transform)
                drift = factor_sum.mult(sig)
                new[i] = libor[i].add(drift.mult(delta)).add(sig.mult(dW))
            libor = new
        value = None
        for i in range(NP - 1, exercise - 1, -1):                                      # Swaption.getValue backward recursion
            payoff = at_exercise[i].sub(strike).mult(delta)
            value = payoff if value is None else value.add(payoff)
            value = value.discount(at_exercise[i], delta)
        return value.floor(0.0).div(numeraire[exercise]).average()

    fac = fc.RandomVariableDifferentiableAADFactory(plain)
    sig = fac.createRandomVariable(0.0, sigma0)
    k0 = fc.stats()["n_tape_kernels"]
    V = swaption_value(sig)
    vega = V.getGradient()[sig.getID()].getAverage()
    kernels = fc.stats()["n_tape_kernels"] - k0
    h = 1e-5
    up = swaption_value(fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, sigma0 + h))).doubleValue()
    dn = swaption_value(fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, sigma0 - h))).doubleValue()
    fd = (up - dn) / (2 * h)
    assert V.doubleValue() > 0 and vega > 0
    assert abs(vega - fd) <= 2e-3 * abs(fd), (vega, fd)
    # Bachelier: value = annuity * sigma * sqrt(T / 2 pi) for the ATM option -> vega ~ value / sigma
    assert abs(vega - V.doubleValue() / sigma0) <= 0.15 * vega
    assert kernels < 60, f"primal + adjoint sweep should fuse into few launches, used {kernels}"


def _aad_chain(X, Y):
    u = X.mult(Y).add(X.mult(0.3).exp()).sub(Y.log().div(X.mult(0.5).add(1.0)))
    v = u.squared().add(1.0).sqrt().accrue(Y, 0.25)
    w = v.sub(2.0).floor(0.0).add(X.cap(1.0)).add(Y.invert()).add(X.discount(Y, 0.5)).add(X.abs().pow(1.5))
    return X.sub(1.0).choose(w.mult(X), w.addProduct(Y, Y))


def test_aad_on_the_gpu_matches_aad_on_the_oracle(fc, O):
    """The SAME RandomVariableDifferentiableAAD code over RandomVariableCuda and over the CPU oracle twin (RandomVariableFromFloatArray
    semantics, oracle/oracle_random_variable.py): primal values bit-equal where every op is exact, every gradient within 1e-5
    relative — an oracle-backed check of the reverse sweep instead of finite differences of the GPU computation itself."""
    from oracle.oracle_random_variable import OracleRandomVariable
    n = 40_000
    rng = np.random.default_rng(11)
    x0, y0 = rng.random(n) + 0.5, rng.random(n) + 0.5
    res = {}
    for name, ctor in (("gpu", fc.RandomVariableCuda), ("cpu", OracleRandomVariable)):
        X, Y = fc.RandomVariableDifferentiableAAD(ctor(0.0, x0)), fc.RandomVariableDifferentiableAAD(ctor(0.0, y0))
        F = _aad_chain(X, Y)
        g = F.getGradient()
        mean = F.average()
        gm = mean.getGradient()
        res[name] = (F.getRealizations(), g[X.getID()].getRealizations(), g[Y.getID()].getRealizations(),
                     gm[X.getID()].getRealizations(), mean.doubleValue())
    (fg, gxg, gyg, gmg, mg), (fcpu, gxc, gyc, gmc, mc) = res["gpu"], res["cpu"]
    assert np.max(np.abs(fg - fcpu) / np.maximum(np.abs(fcpu), 1e-6)) <= 2e-7          # exp / log / pow: <= 1 ulp each
    for a, b in ((gxg, gxc), (gyg, gyc), (gmg, gmc)):
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)) <= 1e-5
    assert abs(mg - mc) <= 1e-6 * abs(mc)


def test_aad_retention_policy_bounds_the_device_footprint(fc):
    """a node of the operator tree keeps only what its derivative rule reads, so primal intermediates nobody needs are
    fused away instead of stored; release=True frees retained values during the sweep. Same gradient, a fraction of the memory."""
    import finmath_cuda.differentiable as D
    n, NP, delta, L0, sigma0, strike = 200_000, 12, 0.5, 0.02, 0.006, 0.02
    td = fc.TimeDiscretization(0.0, NP, delta)
    bm = fc.BrownianMotionCuda(td, 1, n, 31415)
    plain = fc.RandomVariableCudaFactory()
    for t in range(NP):
        bm.getBrownianIncrement(t, 0)

    def vega(retention, release):
        D.RETENTION = retention
        try:
            fc.sync(); fc.pool_trim(); fc.reset_stats()
            base = fc.stats()["bytes_in_use"]
            sig = fc.RandomVariableDifferentiableAAD(plain.createRandomVariable(0.0, sigma0))
            libor = [plain.createRandomVariable(0.0, L0) for _ in range(NP)]
            exercise, at_ex = NP // 2, None
            for t in range(NP):
                if t == exercise: at_ex = list(libor)
                dW, acc, new = bm.getBrownianIncrement(t, 0), None, list(libor)
                for i in range(t + 1, NP):
                    tr = sig.mult(libor[i].mult(delta).add(1.0).invert().mult(delta))
                    acc = tr if acc is None else acc.add(tr)
                    new[i] = libor[i].add(acc.mult(sig).mult(delta)).add(sig.mult(dW))
                libor = new
            value = None
            for i in range(NP - 1, exercise - 1, -1):
                payoff = at_ex[i].sub(strike).mult(delta)
                value = payoff if value is None else value.add(payoff)
                value = value.discount(at_ex[i], delta)
            V = value.floor(0.0).average()
            del libor, new, at_ex, value, acc, tr, payoff
            g = V.getGradient(release=release)[sig.getID()].getAverage()
            return g, fc.stats()["bytes_high_water"] - base
        finally:
            D.RETENTION = "needed"

    g_all, mem_all = vega("all", False)
    g_need, mem_need = vega("needed", False)
    g_rel, mem_rel = vega("needed", True)
    assert g_all > 0 and abs(g_need - g_all) <= 1e-9 * abs(g_all) and abs(g_rel - g_all) <= 1e-9 * abs(g_all)
    assert mem_need <= 0.6 * mem_all, (mem_need, mem_all)
    assert mem_rel <= mem_need, (mem_rel, mem_need)
