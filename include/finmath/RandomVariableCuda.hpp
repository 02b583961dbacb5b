// RandomVariableCuda.hpp — the product: RandomVariableCuda / RandomVariableCudaFactory / BrownianMotionCuda in C++
// over the C ABI of include/fmcuda.h. Same public surface as
//   net.finmath.cuda.montecarlo.RandomVariableCuda (type priority 20, RVC:568),
//   net.finmath.cuda.montecarlo.RandomVariableCudaFactory (RandomVariableCudaFactory.java:18-35),
//   BrownianMotionCudaWithRandomVariableCuda's BrownianMotion shape (BMC:79-250) with the Mersenne stream.
// Errors of the runtime surface as C++ exceptions (CudaException <-> JCuda's, OutOfMemoryError <-> RVC:373-376).
#pragma once
#include <mutex>
#include <new>
#include <string>

#include "../fmcuda.h"
#include "RandomVariableImpl.hpp"

namespace finmath {

struct CudaException : std::runtime_error { using std::runtime_error::runtime_error; };

inline void fmc_check(int status) {
    if (status == FMC_OK) return;
    const std::string msg = fmc_last_error();
    if (status == FMC_ERR_OOM) throw std::bad_alloc();
    if (status == FMC_ERR_SIZE) throw std::out_of_range(msg);
    if (status == FMC_ERR_INVALID) throw std::invalid_argument(msg);
    if (status == FMC_ERR_UNSUPPORTED) throw UnsupportedOperationException(msg);
    throw CudaException("[" + std::to_string(status) + "] " + msg);
}

struct CudaBackend {
    static constexpr int kTypePriority = 20;                       // RVC:568
    struct Vec { fmc_vec h = 0; Vec() = default; explicit Vec(fmc_vec x) : h(x) {}
                 Vec(Vec&& o) noexcept : h(o.h) { o.h = 0; } Vec& operator=(Vec&& o) noexcept { h = o.h; o.h = 0; return *this; }
                 Vec(const Vec&) = delete; Vec& operator=(const Vec&) = delete; };
    static void release(Vec& v) { if (v.h) { fmc_vec_release(v.h); v.h = 0; } }
    static Vec from_f64(const double* p, int64_t n) { fmc_vec h; fmc_check(fmc_vec_from_f64(p, n, &h)); return Vec(h); }
    static Vec vs(int op, const Vec& a, double s, int64_t) { fmc_vec h; fmc_check(fmc_op_vs(op, a.h, s, &h)); return Vec(h); }
    static Vec v(int op, const Vec& a, int64_t) { fmc_vec h; fmc_check(fmc_op_v(op, a.h, &h)); return Vec(h); }
    static Vec vv(int op, const Vec& a, const Vec& b, int64_t, int64_t) { fmc_vec h; fmc_check(fmc_op_vv(op, a.h, b.h, &h)); return Vec(h); }
    static Vec vvs(int op, const Vec& a, const Vec& b, double s, int64_t) { fmc_vec h; fmc_check(fmc_op_vvs(op, a.h, b.h, s, &h)); return Vec(h); }
    static Vec vvv(int op, const Vec& a, const Vec& b, const Vec& c, int64_t) { fmc_vec h; fmc_check(fmc_op_vvv(op, a.h, b.h, c.h, &h)); return Vec(h); }
    static Vec choose(const Vec& t, const Vec* a, double sa, const Vec* b, double sb, int64_t) {
        fmc_vec h; fmc_check(fmc_op_choose(t.h, a ? a->h : 0, sa, b ? b->h : 0, sb, &h)); return Vec(h);
    }
    static double reduce(int kind, const Vec& a, int64_t, const Vec* w) { double r; fmc_check(fmc_reduce(kind, a.h, w ? w->h : 0, &r)); return r; }
    static double quantile(const Vec& a, int64_t, double q) { double r; fmc_check(fmc_quantile(a.h, q, &r)); return r; }
    static double quantile_expectation(const Vec& a, int64_t, double q0, double q1) { double r; fmc_check(fmc_quantile_expectation(a.h, q0, q1, &r)); return r; }
    static double get(const Vec& a, int64_t, int64_t i) { double r; fmc_check(fmc_vec_get(a.h, i, &r)); return r; }
    static std::vector<double> to_f64(const Vec& a, int64_t n) { std::vector<double> r((size_t)n); fmc_check(fmc_vec_to_f64(a.h, r.data(), n)); return r; }
};

using RandomVariableCuda = RandomVariableImpl<CudaBackend>;

class RandomVariableCudaFactory : public RandomVariableFactory {
public:
    using RandomVariableFactory::createRandomVariable;
    // Extension for callers that keep their doubles in pinned memory (fmc_host_alloc): asynchronous DMA + cast on the device.
    // The array is read after the call returns; see fmc_vec_from_f64_pinned.
    RV createRandomVariableFromPinned(double time, const double* pinnedValues, int64_t n) const {
        fmc_vec h = 0;
        fmc_check(fmc_vec_from_f64_pinned(pinnedValues, n, &h));
        return RandomVariableCuda::of(time, CudaBackend::Vec(h), n);
    }
    RV createRandomVariable(double time, double value) const override { return RandomVariableCuda::of(time, value); }                       // RVCF:26-29
    RV createRandomVariable(double time, const double* values, int64_t n) const override { return RandomVariableCuda::of(time, values, n); }   // RVCF:31-34
};

// fused conditional-expectation regression (see csrc/regression_kernel.cu): XtX (k*k) and XtY (k) in one pass
inline void cudaRegressionNormalEquations(const std::vector<RV>& basis, const RV& y, std::vector<double>& XtX, std::vector<double>& XtY) {
    const int k = (int)basis.size();
    std::vector<std::shared_ptr<const RandomVariableCuda>> keep;
    std::vector<fmc_vec> h((size_t)k); std::vector<double> s((size_t)k);
    for (int i = 0; i < k; i++) {
        auto c = RandomVariableCuda::as_self(basis[(size_t)i]);
        keep.push_back(c);
        h[(size_t)i] = c->isDeterministic() ? 0 : c->vec().h;
        s[(size_t)i] = c->isDeterministic() ? c->doubleValue() : 0.0;
    }
    auto cy = RandomVariableCuda::as_self(y);
    if (cy->isDeterministic()) throw std::invalid_argument("the dependent variable must be stochastic");
    XtX.assign((size_t)k * k, 0.0); XtY.assign((size_t)k, 0.0);
    fmc_check(fmc_regression_normal_eq(h.data(), s.data(), k, cy->vec().h, XtX.data(), XtY.data()));
}

class BrownianMotionCuda : public BrownianMotion {
public:
    // seedMode 0: net.finmath.randomnumbers.MersenneTwister(long) (finmath-lib 5.x); 1: commons-math3 MersenneTwister(int)
    BrownianMotionCuda(TimeDiscretization td, int numberOfFactors, int64_t numberOfPaths, int seed, int seedMode = 0,
                       int64_t p0 = 0, int64_t p1 = -1)
        : td_(std::move(td)), factors_(numberOfFactors), paths_(numberOfPaths), seed_(seed), seedMode_(seedMode),
          p0_(p0), p1_(p1 < 0 ? numberOfPaths : p1) {}
    RV getBrownianIncrement(int timeIndex, int factor) override {                                  // BMC:122-136
        std::lock_guard<std::mutex> lock(mu_);
        if (inc_.empty()) generate();
        return inc_[(size_t)timeIndex * factors_ + factor];
    }
    const TimeDiscretization& getTimeDiscretization() const override { return td_; }
    int getNumberOfFactors() const override { return factors_; }
    int64_t getNumberOfPaths() const override { return paths_; }
    RV getRandomVariableForConstant(double value) const override { return RandomVariableCuda::of(-1.7976931348623157e308, value); }   // BMC:199-202
    std::shared_ptr<BrownianMotionCuda> getCloneWithModifiedSeed(int seed) const {                  // BMC:111-114
        return std::make_shared<BrownianMotionCuda>(td_, factors_, paths_, seed, seedMode_, p0_, p1_);
    }
    std::shared_ptr<BrownianMotionCuda> getCloneWithModifiedTimeDiscretization(TimeDiscretization newTimeDiscretization) const {   // BMC:116-120
        return std::make_shared<BrownianMotionCuda>(std::move(newTimeDiscretization), factors_, paths_, seed_, seedMode_, p0_, p1_);
    }
    int getSeed() const { return seed_; }
private:
    void generate() {                                                                               // BMC:141-182
        const int T = td_.getNumberOfTimeSteps();
        std::vector<double> sq((size_t)T);
        for (int t = 0; t < T; t++) sq[(size_t)t] = std::sqrt(td_.getTimeStep(t));
        std::vector<fmc_vec> h((size_t)T * factors_);
        fmc_check(fmc_brownian_generate(seedMode_, seed_, T, factors_, p0_, p1_, sq.data(), h.data()));
        inc_.resize(h.size());
        for (int t = 0; t < T; t++)
            for (int f = 0; f < factors_; f++)
                inc_[(size_t)t * factors_ + f] = RandomVariableCuda::of(td_.getTime(t + 1), CudaBackend::Vec(h[(size_t)t * factors_ + f]), p1_ - p0_);
    }
    TimeDiscretization td_;
    int factors_; int64_t paths_; int seed_, seedMode_; int64_t p0_, p1_;
    std::vector<RV> inc_;
    std::mutex mu_;
};

}  // namespace finmath
