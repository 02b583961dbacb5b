// driver_api.cpp — C entry points of the workload drivers (benchmarks and whole-workload parity tests).
// Compiled twice from this one source:
//   -DFMD_BACKEND_CUDA   -> lib/libfmdrivers_cuda.so   (RandomVariableCudaFactory + BrownianMotionCuda, links libfmcuda.so)
//   -DFMD_BACKEND_ORACLE -> oracle/libfmdrivers_oracle.so (RandomVariableFloatFactory + CPU Mersenne generator; test / baseline only)
//   -DFMD_BACKEND_ORACLE_F64 -> oracle/libfmdrivers_oracle_f64.so (finmath-lib's default RandomVariableFromArrayFactory on doubles; CPU baseline only)
#include <chrono>
#include <cstring>
#include <string>

#include "workloads.hpp"

#if defined(FMD_BACKEND_CUDA)
#include "../../include/finmath/RandomVariableCuda.hpp"
#elif defined(FMD_BACKEND_ORACLE)
#include "../../oracle/RandomVariableFromFloatArray.hpp"
#elif defined(FMD_BACKEND_ORACLE_F64)
#include "../../oracle/RandomVariableFromFloatArray.hpp"    // BrownianMotionFromMersenneRandomNumbers (factory-agnostic)
#include "../../oracle/RandomVariableFromDoubleArray.hpp"
#else
#error "define FMD_BACKEND_CUDA, FMD_BACKEND_ORACLE or FMD_BACKEND_ORACLE_F64"
#endif

using namespace workloads;

namespace {

thread_local std::string g_err;

FactoryPtr make_factory() {
#if defined(FMD_BACKEND_CUDA)
    return std::make_shared<RandomVariableCudaFactory>();
#elif defined(FMD_BACKEND_ORACLE_F64)
    return std::make_shared<RandomVariableFromArrayFactory>();
#else
    return std::make_shared<RandomVariableFloatFactory>();
#endif
}

std::shared_ptr<BrownianMotion> make_brownian(const TimeDiscretization& td, int factors, int64_t paths, int seed, int seed_mode,
                                              const FactoryPtr& factory, int64_t p0, int64_t p1) {
#if defined(FMD_BACKEND_CUDA)
    (void)factory;
    return std::make_shared<BrownianMotionCuda>(td, factors, paths, seed, seed_mode, p0, p1);
#else
    return std::make_shared<BrownianMotionFromMersenneRandomNumbers>(td, factors, paths, seed, factory, seed_mode, p0, p1);
#endif
}

RegressionFn regression_fn() {
#if defined(FMD_BACKEND_CUDA)
    return cudaRegressionNormalEquations;
#elif defined(FMD_BACKEND_ORACLE_F64)
    return doubleRegressionNormalEquations;
#else
    return oracleRegressionNormalEquations;
#endif
}

// Brownian motion whose increments live in HOST double arrays and enter the backend through
// RandomVariableFactory.createRandomVariable(time, double[]) — exactly what the reference's heavy tests do with
// BrownianMotionFromMersenneRandomNumbers + RandomVariableCudaFactory (T-ATM:283): T*F host->device uploads.
class HostArrayBrownianMotion : public BrownianMotion {
public:
    HostArrayBrownianMotion(BrownianMotion& source, FactoryPtr factory) : td_(source.getTimeDiscretization()),
        factors_(source.getNumberOfFactors()), paths_(source.getNumberOfPaths()), factory_(std::move(factory)) {
        const int T = td_.getNumberOfTimeSteps();
        host_.resize((size_t)T * factors_);
        for (int t = 0; t < T; t++)
            for (int f = 0; f < factors_; f++) {
                host_[(size_t)t * factors_ + f] = source.getBrownianIncrement(t, f)->getRealizations();
                local_n_ = (int64_t)host_[(size_t)t * factors_ + f].size();
            }
        inc_.resize(host_.size());
#if defined(FMD_BACKEND_CUDA)
        // a second copy of the same doubles in pinned memory, for the asynchronous upload (fmc_vec_from_f64_pinned)
        const size_t bytes = sizeof(double) * (size_t)local_n_ * host_.size();
        void* p = nullptr;
        fmc_check(fmc_host_alloc(bytes, &p));
        pinned_ = static_cast<double*>(p);
        for (size_t k = 0; k < host_.size(); k++) std::memcpy(pinned_ + k * (size_t)local_n_, host_[k].data(), sizeof(double) * (size_t)local_n_);
#endif
    }
    ~HostArrayBrownianMotion() override {
#if defined(FMD_BACKEND_CUDA)
        inc_.clear();
        if (pinned_) fmc_host_free(pinned_);
#endif
    }
    void reset() { for (auto& r : inc_) r.reset(); }
    void use_pinned(bool on) { use_pinned_ = on; }
    uint64_t host_bytes() const { return (uint64_t)host_.size() * (uint64_t)local_n_ * sizeof(double); }
    RV getBrownianIncrement(int t, int f) override {
        RV& r = inc_[(size_t)t * factors_ + f];
        if (!r) {
#if defined(FMD_BACKEND_CUDA)
            if (use_pinned_) {
                r = static_cast<const RandomVariableCudaFactory&>(*factory_).createRandomVariableFromPinned(
                        td_.getTime(t + 1), pinned_ + ((size_t)t * factors_ + f) * (size_t)local_n_, local_n_);
                return r;
            }
#endif
            r = factory_->createRandomVariable(td_.getTime(t + 1), host_[(size_t)t * factors_ + f].data(), local_n_);
        }
        return r;
    }
    const TimeDiscretization& getTimeDiscretization() const override { return td_; }
    int getNumberOfFactors() const override { return factors_; }
    int64_t getNumberOfPaths() const override { return paths_; }
    RV getRandomVariableForConstant(double v) const override { return factory_->createRandomVariable(v); }
private:
    TimeDiscretization td_;
    int factors_; int64_t paths_, local_n_ = 0;
    FactoryPtr factory_;
    std::vector<std::vector<double>> host_;
    std::vector<RV> inc_;
    double* pinned_ = nullptr;
    bool use_pinned_ = false;
};

struct LmmHandle {
    FactoryPtr factory;
    std::shared_ptr<BrownianMotion> device_brownian;            // generated by the backend (device resident for CUDA)
    std::shared_ptr<HostArrayBrownianMotion> host_brownian;     // host arrays, re-uploaded every step (end-to-end arm)
    std::vector<double> L0;
    double delta;
    std::vector<SwaptionSpec> products;
    PiecewiseConstantVolatility vol;
    std::unique_ptr<LIBORMarketModel> last_model;               // keeps the last simulation alive (Bermudan on top of it)
    int valuation_threads = 1;                                  // host threads valuing the calibration products (T-ATM:319: 1)
    bool price_products = false;                                // workloads.hpp: lmm_value_products
};

template <typename F>
int guarded(F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown error"; return -1; }
}

}  // namespace

extern "C" {

const char* fmd_last_error(void) { return g_err.c_str(); }

int fmd_backend(void) {
#if defined(FMD_BACKEND_CUDA)
    return 1;
#elif defined(FMD_BACKEND_ORACLE_F64)
    return 2;
#else
    return 0;
#endif
}

// Black-Scholes call by Euler Monte-Carlo (MonteCarloBlackScholesModelTest.java)
int fmd_bs_call(int64_t n_paths, int n_steps, double dt, int seed, int seed_mode, double S0, double r, double sigma,
                double maturity, double strike, double* value, double* analytic) {
    return guarded([&] {
        FactoryPtr fac = make_factory();
        TimeDiscretization td(0.0, n_steps, dt);
        auto bm = make_brownian(td, 1, n_paths, seed, seed_mode, fac, 0, n_paths);
        const BlackScholesResult res = black_scholes_call(fac, *bm, S0, r, sigma, maturity, strike);
        *value = res.value; *analytic = res.analytic;
    });
}

// LIBOR market model workload (LIBORMarketModelCalibrationATMTest.java). Paths [p0,p1) of an n_paths motion.
void* fmd_lmm_create(int64_t n_paths, int n_periods, double delta, int n_factors, int seed, int seed_mode, int64_t p0, int64_t p1) {
    LmmHandle* h = nullptr;
    const int rc = guarded([&] {
        h = new LmmHandle();
        h->factory = make_factory();
        h->delta = delta;
        TimeDiscretization td(0.0, n_periods, delta);
        h->device_brownian = make_brownian(td, n_factors, n_paths, seed, seed_mode, h->factory, p0, p1 < 0 ? n_paths : p1);
        h->L0 = synthetic_forward_rates(n_periods, delta);
        h->products = atm_calibration_products(h->L0, delta, n_periods);
    });
    if (rc != 0) { delete h; return nullptr; }
    return h;
}
void fmd_lmm_destroy(void* handle) { delete static_cast<LmmHandle*>(handle); }
int fmd_lmm_num_products(void* handle) { return (int)static_cast<LmmHandle*>(handle)->products.size(); }
int fmd_lmm_num_parameters(void* handle) { return (int)static_cast<LmmHandle*>(handle)->vol.param.size(); }
int fmd_lmm_get_parameters(void* handle, double* out) {
    auto* h = static_cast<LmmHandle*>(handle);
    std::memcpy(out, h->vol.param.data(), sizeof(double) * h->vol.param.size());
    return 0;
}
int fmd_lmm_product_info(void* handle, int k, int* exercise_index, int* n_periods, double* strike, double* target_vol) {
    auto* h = static_cast<LmmHandle*>(handle);
    const SwaptionSpec& s = h->products[(size_t)k];
    *exercise_index = s.exerciseIndex; *n_periods = s.numberOfPeriods; *strike = s.strike; *target_vol = s.targetVolatility;
    return 0;
}
// capture the increments into host arrays (outside any timed region) for the end-to-end arm
int fmd_lmm_prepare_host_brownian(void* handle, uint64_t* host_bytes) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        if (!h->host_brownian) h->host_brownian = std::make_shared<HostArrayBrownianMotion>(*h->device_brownian, h->factory);
        if (host_bytes) *host_bytes = h->host_brownian->host_bytes();
    });
}
// one pass of the calibration inner loop: Euler simulation (n_periods steps x n_periods rates) + valuation of all
// calibration swaptions. from_host != 0: the Brownian increments are uploaded from host arrays inside this call.
int fmd_lmm_step(void* handle, const double* vol_params, int from_host, double* values_out) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        if (vol_params) std::memcpy(h->vol.param.data(), vol_params, sizeof(double) * h->vol.param.size());
        std::shared_ptr<BrownianMotion> bm = h->device_brownian;
        if (from_host) {
            if (!h->host_brownian) throw std::runtime_error("call fmd_lmm_prepare_host_brownian first");
            h->last_model.reset();                 // drop the previous simulation before its inputs
            h->host_brownian->reset();
            h->host_brownian->use_pinned(from_host == 2);   // 1: pageable double[] (the reference API); 2: pinned, asynchronous
            bm = h->host_brownian;
        }
        h->last_model.reset(new LIBORMarketModel(h->factory, bm, h->L0, h->vol));
        const std::vector<double> v = lmm_value_products(*h->last_model, h->products, h->valuation_threads, h->price_products);
        std::memcpy(values_out, v.data(), sizeof(double) * v.size());
    });
}
int fmd_lmm_set_valuation_threads(void* handle, int threads) {
    return guarded([&] {
        if (threads < 1 || threads > 64) throw std::runtime_error("valuation threads must be in 1..64");
        static_cast<LmmHandle*>(handle)->valuation_threads = threads;
    });
}
int fmd_lmm_set_price_products(void* handle, int on) {
    return guarded([&] { static_cast<LmmHandle*>(handle)->price_products = on != 0; });
}
// selected simulated rates for parity checks: L_i(t) realisations
int fmd_lmm_get_libor(void* handle, int time_index, int libor_index, double* out, int64_t n) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        if (!h->last_model) throw std::runtime_error("no simulation");
        const std::vector<double> r = h->last_model->getLIBOR(time_index, libor_index)->getRealizations();
        if ((int64_t)r.size() != n) throw std::runtime_error("size mismatch");
        std::memcpy(out, r.data(), sizeof(double) * (size_t)n);
    });
}
// implied normal volatilities of the calibration products from their values (host scalar math)
int fmd_lmm_implied_vols(void* handle, const double* values, double* vols_out) {
    auto* h = static_cast<LmmHandle*>(handle);
    Curve curve(h->L0, h->delta);
    for (size_t k = 0; k < h->products.size(); k++) {
        const SwaptionSpec& s = h->products[k];
        vols_out[k] = atm_normal_implied_vol(values[k], curve.annuity(s.exerciseIndex, s.numberOfPeriods, h->delta), s.exerciseIndex * h->delta);
    }
    return 0;
}
// Bermudan swaption on the last simulation (simulates first if needed)
int fmd_lmm_bermudan(void* handle, int first_exercise, int last_exercise, int stride, int swap_end, double strike, double* value) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        if (!h->last_model) {
            h->last_model.reset(new LIBORMarketModel(h->factory, h->device_brownian, h->L0, h->vol));
            h->last_model->simulate();
        }
        BermudanSpec spec{first_exercise, last_exercise, stride, swap_end, strike};
        *value = bermudan_swaption_value(*h->last_model, spec, regression_fn());
    });
}
// initial forward rates and calibration products from the market curve of LIBORMarketModelCalibrationATMTest.java:526-663 instead of the synthetic one
int fmd_lmm_use_market_curve(void* handle) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        const int n = (int)h->L0.size();
        h->L0 = MarketCurveATM().forwardRates(n, h->delta);
        h->products = atm_calibration_products(h->L0, h->delta, n);
        h->last_model.reset();
    });
}
int fmd_lmm_get_forward_rates(void* handle, double* out) {
    auto* h = static_cast<LmmHandle*>(handle);
    std::memcpy(out, h->L0.data(), sizeof(double) * h->L0.size());
    return 0;
}
// The calibration of LIBORMarketModelCalibrationATMTest.java:317-358: Levenberg-Marquardt over the piecewise-constant volatility
// parameters, objective = implied normal volatilities of the calibration swaptions (SwaptionSimple, VOLATILITYNORMAL) against
// their market quotes; every evaluation is one simulation + valuation (fmd_lmm_step). params_out: calibrated parameters;
// info_out[6]: iterations, evaluations, root mean squared error, mean deviation (model - target volatility), seconds, seconds per evaluation.
int fmd_lmm_calibrate(void* handle, int max_iterations, double accuracy, double lambda, double parameter_step, double* params_out, double* info_out) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        const size_t nprod = h->products.size();
        std::vector<double> target(nprod), annuity(nprod), maturity(nprod);
        Curve curve(h->L0, h->delta);
        for (size_t k = 0; k < nprod; k++) {
            const SwaptionSpec& sp = h->products[k];
            target[k] = sp.targetVolatility;
            annuity[k] = curve.annuity(sp.exerciseIndex, sp.numberOfPeriods, h->delta);
            maturity[k] = sp.exerciseIndex * h->delta;
        }
        const auto t0 = std::chrono::steady_clock::now();
        auto objective = [&](const std::vector<double>& p) {
            h->last_model.reset();
            PiecewiseConstantVolatility vol = h->vol;
            vol.param = p;
            h->last_model.reset(new LIBORMarketModel(h->factory, h->device_brownian, h->L0, vol));
            std::vector<double> v = lmm_value_products(*h->last_model, h->products, h->valuation_threads, h->price_products);
            for (size_t k = 0; k < nprod; k++) v[k] = atm_normal_implied_vol(v[k], annuity[k], maturity[k]);
            return v;
        };
        const LevenbergMarquardtResult r = levenberg_marquardt(objective, h->vol.param, target, max_iterations, accuracy, lambda, parameter_step);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        h->vol.param = r.parameters;
        std::memcpy(params_out, r.parameters.data(), sizeof(double) * r.parameters.size());
        double dev = 0.0;
        for (size_t k = 0; k < nprod; k++) dev += r.values[k] - target[k];
        info_out[0] = r.iterations; info_out[1] = r.evaluations; info_out[2] = r.rootMeanSquaredError; info_out[3] = dev / (double)nprod;
        info_out[4] = secs; info_out[5] = secs / std::max(1, r.evaluations);
    });
}
int fmd_lmm_set_parameters(void* handle, const double* params) {
    auto* h = static_cast<LmmHandle*>(handle);
    std::memcpy(h->vol.param.data(), params, sizeof(double) * h->vol.param.size());
    return 0;
}
int fmd_lmm_simulate(void* handle) {
    return guarded([&] {
        auto* h = static_cast<LmmHandle*>(handle);
        h->last_model.reset(new LIBORMarketModel(h->factory, h->device_brownian, h->L0, h->vol));
        h->last_model->simulate();
    });
}

}  // extern "C"
