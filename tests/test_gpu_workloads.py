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
    # model reproduces its own flat 0.5% volatility to MC accuracy
    iv = mg.implied_vols(mg.step(mg.parameters() * 0 + 0.005))
    assert np.all(np.abs(iv - 0.005) < 0.0015)


def test_bermudan_swaption_matches_oracle(libs):
    """BASELINE config 3 (small): backward induction with conditional-expectation regression and choose()."""
    gpu, cpu = libs
    mg, mc = gpu.lmm(20_000), cpu.lmm(20_000)
    mg.simulate(); mc.simulate()
    for spec in ((10, 30, 2, 40, 0.02), (4, 20, 4, 24, 0.015)):
        vg, vc = mg.bermudan(*spec), mc.bermudan(*spec)
        assert vg > 0
        assert abs(vg - vc) <= 1e-4 * abs(vc), (spec, vg, vc)
