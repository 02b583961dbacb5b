"""Whole-workload parity on the GPU: the same driver source (drivers/workloads.hpp) runs on RandomVariableCuda
(product, C ABI) and on the CPU oracle twin; results must agree within the north-star tolerances
(prices 1e-4 relative; here far tighter because every elementwise op is bit-exact)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs(fc):
    from finmath_cuda.workloads import DriverLib
    from oracle.workloads_oracle import driver
    return DriverLib(), driver()


def test_black_scholes_call_config1(libs):
    """BASELINE config 1 / MonteCarloBlackScholesModelTest.java:62-76,156: 100k paths x 100 steps, |MC - analytic| < 0.005."""
    gpu, cpu = libs
    vg, ag = gpu.bs_call(100_000)
    vc, ac = cpu.bs_call(100_000)
    assert ag == ac and abs(ag - 0.18993678) < 1e-7
    assert abs(vg - ag) < 0.005
    assert abs(vg - vc) <= 1e-4 * abs(vc)
    assert abs(vg - vc) <= 1e-9 * abs(vc), (vg, vc)


def test_black_scholes_call_1m_paths(libs):
    gpu, _ = libs
    v, a = gpu.bs_call(1_000_000)
    assert abs(v - a) < 0.005            # MonteCarloBlackScholesModelTest.java:156 at the test's own size


@pytest.mark.parametrize("paths", [1000, 4099])
def test_lmm_simulation_and_swaptions_match_oracle(libs, paths):
    gpu, cpu = libs
    mg, mc = gpu.lmm(paths), cpu.lmm(paths)
    assert mg.n_products == mc.n_products == 144 and mg.n_parameters == 48
    vg, vc = mg.step(), mc.step()
    # NORMAL state space: only + - * / -> the simulated rates are bit-identical
    for (t, i) in ((1, 1), (10, 40), (40, 40), (79, 79), (80, 79)):
        assert np.array_equal(mg.libor(t, i), mc.libor(t, i)), (t, i)
    assert np.allclose(vg, vc, rtol=1e-9, atol=1e-15)
    # modified volatility parameters (what the optimiser does between simulations)
    p = mg.parameters() * np.linspace(0.8, 1.3, mg.n_parameters)
    vg_p = mg.step(p)
    assert np.allclose(vg_p, mc.step(p), rtol=1e-9, atol=1e-15)
    # host-array (end-to-end) arm gives the same numbers as the device-resident arm
    mg.prepare_host_brownian()
    assert np.array_equal(mg.step(p, from_host=True), mg.step(p))
    # ... and so does the asynchronous upload from pinned host doubles (fmc_vec_from_f64_pinned, cast on the device)
    assert np.array_equal(mg.step(p, from_host=2), mg.step(p))
    # products valued by three host threads (the library's numberOfThreads > 1): the same numbers
    mg.set_valuation_threads(3)
    assert np.array_equal(mg.step(p), vg_p)
    mg.set_valuation_threads(1)
    # price products: all value vectors first, the averages in a second loop (the runtime sums them in batches) — the same
    # values, a handful of reduction launches instead of one per product
    mg.set_price_products(True)
    import finmath_cuda as fcm
    k0 = fcm.stats()["n_kernels"]
    v2 = mg.step(p)
    k1 = fcm.stats()["n_kernels"]
    mg.set_price_products(False)
    assert np.allclose(v2, vg_p, rtol=1e-12, atol=0)
    mg.step(p)
    assert k1 - k0 < fcm.stats()["n_kernels"] - k1, "batched averages must launch fewer kernels than one reduction per product"
    # model reproduces its own flat 0.5% volatility to MC accuracy
    iv = mg.implied_vols(mg.step(mg.parameters() * 0 + 0.005))
    assert np.all(np.abs(iv - 0.005) < 0.0015)


@pytest.mark.parametrize("paths", [700, 70_000])
def test_windows_of_time_steps_give_the_same_simulation(libs, paths):
    """Option window_levels (several Euler time steps per launch, emitted component by component with the running sums of the
    window's steps kept on chip): same arithmetic per path in a different schedule — the LIBORs are bit-identical
    to one launch per time step, with fewer launches; every geometry; the oracle agrees."""
    import finmath_cuda as fcm
    gpu, cpu = libs
    probes = ((1, 5), (2, 2), (7, 7), (7, 8), (8, 79), (33, 34), (40, 60), (41, 41), (79, 79), (80, 79))

    def run(window, elems):
        fcm.set_option("window_levels", window)
        fcm.set_option("tape_elems", elems)
        m = gpu.lmm(paths)
        k0 = fcm.stats()["n_kernels"]
        m.simulate()
        fcm.sync()
        k1 = fcm.stats()["n_kernels"]
        vals = m.step()
        out = [m.libor(t, i).copy() for (t, i) in probes]
        m.close()
        return out, k1 - k0, vals
    try:
        ref, launches0, vals0 = run(0, 0)
        mc = cpu.lmm(paths)
        mc.step()
        for (t, i), a in zip(probes, ref):
            assert np.array_equal(a, mc.libor(t, i)), (t, i)
        for window, elems in ((1, 0), (2, 0), (3, 0), (4, 16), (3, 8), (6, 8), (2, 4), (5, 4)):
            got, launches, vals = run(window, elems)
            for (t, i), a, b in zip(probes, got, ref):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (window, elems, t, i)
            assert np.allclose(vals, vals0, rtol=1e-12, atol=0), (window, elems)
            if window >= 2:
                assert launches < launches0, (window, elems, launches, launches0)
        # no automatic flush: the whole simulation is still pending when the first swaption is valued and goes through the
        # windows there (Runtime::reduce)
        fcm.set_option("window_levels", 3)
        fcm.set_option("tape_elems", 0)
        fcm.set_option("flush_threshold", 10_000_000)
        fcm.set_option("window_reduce_min", 0)
        m = gpu.lmm(paths)
        vals = m.step()
        for (t, i), b in zip(probes, ref):
            assert np.array_equal(m.libor(t, i).view(np.uint32), b.view(np.uint32)), ("flushed by the first valuation", t, i)
        assert np.allclose(vals, vals0, rtol=1e-12, atol=0)
        m.close()
    finally:
        fcm.set_option("window_levels", 3)
        fcm.set_option("tape_elems", 0)
        fcm.set_option("flush_threshold", 4096)
        fcm.set_option("window_reduce_min", 2048)


def test_windows_adapt_to_a_three_factor_model(libs):
    """Three running sums and three Brownian increments per time step do not fit the register file for windows of three levels:
    the window size adapts (Runtime::run_windows) — LIBORs bit-identical to one launch per time step and to the oracle, and about
    as many launches."""
    import finmath_cuda as fcm
    gpu, cpu = libs
    paths = 5000
    probes = ((1, 5), (7, 8), (20, 30), (39, 39), (40, 39))

    def run(window):
        fcm.set_option("window_levels", window)
        m = gpu.lmm(paths, 40, 0.5, 3)
        for _ in range(2):
            m.simulate()
            fcm.sync()
        k0 = fcm.stats()["n_kernels"]
        m.simulate()
        fcm.sync()
        launches = fcm.stats()["n_kernels"] - k0
        out = [m.libor(t, i).copy() for (t, i) in probes]
        m.close()
        return launches, out
    try:
        l0, ref = run(0)
        l3, got = run(3)
        mc = cpu.lmm(paths, 40, 0.5, 3)
        mc.simulate()
        for (t, i), a, b in zip(probes, got, ref):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (t, i)
            assert np.array_equal(a, mc.libor(t, i)), (t, i)
        assert l3 <= l0 + 8, (l3, l0)
    finally:
        fcm.set_option("window_levels", 3)


def test_bermudan_swaption_matches_oracle(libs):
    """BASELINE config 3 (small): backward induction with conditional-expectation regression and choose()."""
    gpu, cpu = libs
    mg, mc = gpu.lmm(20_000), cpu.lmm(20_000)
    mg.simulate(); mc.simulate()
    for spec in ((10, 30, 2, 40, 0.02), (4, 20, 4, 24, 0.015)):
        vg, vc = mg.bermudan(*spec), mc.bermudan(*spec)
        assert vg > 0
        assert abs(vg - vc) <= 1e-4 * abs(vc), (spec, vg, vc)


def test_lmm_parity_at_100k_paths_and_bermudan_at_200k(libs):
    """The headline workload at sizes nearer the bench's (parity was only shown at 1 000 / 4 099 / 20 000 paths)."""
    gpu, cpu = libs
    mg, mc = gpu.lmm(100_000), cpu.lmm(100_000)
    vg, vc = mg.step(), mc.step()
    assert np.array_equal(mg.libor(80, 79), mc.libor(80, 79)) and np.array_equal(mg.libor(37, 52), mc.libor(37, 52))
    assert np.max(np.abs(vg - vc) / np.abs(vc)) <= 1e-9
    del mg, mc
    mg, mc = gpu.lmm(200_000), cpu.lmm(200_000)
    mg.simulate(); mc.simulate()
    spec = (10, 30, 2, 40, 0.02)
    vg, vc = mg.bermudan(*spec), mc.bermudan(*spec)
    assert abs(vg - vc) <= 1e-4 * abs(vc), (vg, vc)


def test_lmm_atm_calibration_matches_the_oracle_and_the_reference_bound(libs):
    """BASELINE.json metric "LMM ATM calibration": LIBORMarketModelCalibrationATMTest.java:154-358 — 10 000 paths, one factor, seed
    31415, Levenberg-Marquardt (lambda 0.1, accuracy 1e-7, parameter step 1e-4, one thread) over the 48 piecewise-constant
    volatilities against the 144 in-horizon ATM swaption quotes, initial rates from the EUR swap curve of the test.
    (a) the same optimiser source on the CUDA backend and on the CPU oracle, for a fixed budget of iterations: calibrated
        parameters and mean deviation within 1e-4 relative (north star);
    (b) the CUDA run carried to convergence: |mean deviation| < 2e-4, the reference's own pass criterion (T-ATM:466)."""
    gpu, cpu = libs
    mg, mc = gpu.lmm(10_000), cpu.lmm(10_000)
    mg.use_market_curve(); mc.use_market_curve()
    assert np.array_equal(mg.forward_rates(), mc.forward_rates()) and mg.n_products == mc.n_products == 144
    rg, rc = mg.calibrate(max_iterations=2), mc.calibrate(max_iterations=2)
    assert rg["iterations"] == rc["iterations"] and rg["evaluations"] == rc["evaluations"]
    assert np.max(np.abs(rg["parameters"] - rc["parameters"]) / np.abs(rc["parameters"])) <= 1e-4
    assert abs(rg["mean_deviation"] - rc["mean_deviation"]) <= 1e-4 * max(abs(rc["mean_deviation"]), 1e-4)
    assert abs(rg["rms_error"] - rc["rms_error"]) <= 1e-4 * rc["rms_error"]
    full = mg.calibrate(max_iterations=200)          # continues from the parameters reached above
    assert abs(full["mean_deviation"]) < 2e-4, full
    assert full["rms_error"] < 2e-4, full
