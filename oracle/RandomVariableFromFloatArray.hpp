// RandomVariableFromFloatArray.hpp — CPU ORACLE twin in C++ (test infrastructure, NOT product code).
//
// C++ object model of the reference's CPU float vector
// (/root/reference/src/main/java/net/finmath/cuda/cpu/montecarlo/RandomVariableFromFloatArray.java, type priority 1,
// RVF:47) built on the arithmetic of oracle/fm_oracle.c, plus RandomVariableFloatFactory (RandomVariableFloatFactory.java)
// and BrownianMotionFromMersenneRandomNumbers (finmath-lib, restated in fm_oracle.c:orc_brownian). Used by the oracle
// build of the workload drivers: parity checks of whole workloads and the CPU baseline of bench.py.
#pragma once
#include <mutex>
#include <vector>

#include "../include/finmath/RandomVariableImpl.hpp"
#include "fm_oracle.h"

namespace finmath {

struct FloatArrayBackend {
    static constexpr int kTypePriority = 1;                        // RVF:47
    using Vec = std::vector<float>;
    static void release(Vec&) {}
    static Vec from_f64(const double* p, int64_t n) { Vec v((size_t)n); orc_from_f64(p, v.data(), n); return v; }
    static Vec vs(int op, const Vec& a, double s, int64_t n) { Vec r((size_t)n); orc_op_vs(op, a.data(), s, r.data(), n); return r; }
    static Vec v(int op, const Vec& a, int64_t n) { Vec r((size_t)n); orc_op_v(op, a.data(), r.data(), n); return r; }
    static Vec vv(int op, const Vec& a, const Vec& b, int64_t n, int64_t nb) {
        if (n != nb) throw std::out_of_range("operand sizes differ");
        Vec r((size_t)n); orc_op_vv(op, a.data(), b.data(), r.data(), n); return r;
    }
    static Vec vvs(int op, const Vec& a, const Vec& b, double s, int64_t n) { Vec r((size_t)n); orc_op_vvs(op, a.data(), b.data(), s, r.data(), n); return r; }
    static Vec vvv(int op, const Vec& a, const Vec& b, const Vec& c, int64_t n) { Vec r((size_t)n); orc_op_vvv(op, a.data(), b.data(), c.data(), r.data(), n); return r; }
    static Vec choose(const Vec& t, const Vec* a, double sa, const Vec* b, double sb, int64_t n) {           // RVF:1277-1284
        Vec r((size_t)n);
        const float fa = (float)sa, fb = (float)sb;
        for (int64_t i = 0; i < n; i++) r[(size_t)i] = ((double)t[(size_t)i] >= 0.0) ? (a ? (*a)[(size_t)i] : fa) : (b ? (*b)[(size_t)i] : fb);
        return r;
    }
    static double reduce(int kind, const Vec& a, int64_t n, const Vec* w) {
        switch (kind) {
        case R_SUM: return orc_average(a.data(), n) * (double)n;
        case R_AVERAGE: return orc_average(a.data(), n);
        case R_VARIANCE: return orc_variance(a.data(), n);
        case R_SAMPLE_VARIANCE: return orc_sample_variance(a.data(), n);
        case R_MIN: return orc_min(a.data(), n);
        case R_MAX: return orc_max(a.data(), n);
        case R_AVERAGE_W: return orc_average_w(a.data(), w->data(), n);
        case R_VARIANCE_W: return orc_variance_w(a.data(), w->data(), n);
        default: throw std::invalid_argument("bad reduction kind");
        }
    }
    static double quantile(const Vec& a, int64_t n, double q) { return orc_quantile(a.data(), n, q); }
    static double quantile_expectation(const Vec& a, int64_t n, double q0, double q1) { return orc_quantile_expectation(a.data(), n, q0, q1); }
    static double get(const Vec& a, int64_t n, int64_t i) { if (i < 0 || i >= n) throw std::out_of_range("index"); return a[(size_t)i]; }
    static std::vector<double> to_f64(const Vec& a, int64_t n) { std::vector<double> r((size_t)n); orc_to_f64(a.data(), r.data(), n); return r; }
};

using RandomVariableFromFloatArray = RandomVariableImpl<FloatArrayBackend>;

class RandomVariableFloatFactory : public RandomVariableFactory {
public:
    using RandomVariableFactory::createRandomVariable;
    RV createRandomVariable(double time, double value) const override { return RandomVariableFromFloatArray::of(time, value); }
    RV createRandomVariable(double time, const double* values, int64_t n) const override { return RandomVariableFromFloatArray::of(time, values, n); }
};

inline void oracleRegressionNormalEquations(const std::vector<RV>& basis, const RV& y, std::vector<double>& XtX, std::vector<double>& XtY) {
    const int k = (int)basis.size();
    std::vector<std::shared_ptr<const RandomVariableFromFloatArray>> keep;
    std::vector<const float*> p((size_t)k); std::vector<double> s((size_t)k);
    for (int i = 0; i < k; i++) {
        auto c = RandomVariableFromFloatArray::as_self(basis[(size_t)i]);
        keep.push_back(c);
        p[(size_t)i] = c->isDeterministic() ? nullptr : c->vec().data();
        s[(size_t)i] = c->isDeterministic() ? c->doubleValue() : 0.0;
    }
    auto cy = RandomVariableFromFloatArray::as_self(y);
    XtX.assign((size_t)k * k, 0.0); XtY.assign((size_t)k, 0.0);
    orc_regression_normal_eq(p.data(), s.data(), k, cy->vec().data(), cy->size(), XtX.data(), XtY.data());
}

// finmath-lib BrownianMotionFromMersenneRandomNumbers: CPU generation, wrapped through ANY RandomVariableFactory
// (LIBORMarketModelCalibrationATMTest.java:283 passes RandomVariableCudaFactory -> T*F uploads of host doubles).
class BrownianMotionFromMersenneRandomNumbers : public BrownianMotion {
public:
    BrownianMotionFromMersenneRandomNumbers(TimeDiscretization td, int numberOfFactors, int64_t numberOfPaths, int seed,
                                            std::shared_ptr<const RandomVariableFactory> factory, int seedMode = 0,
                                            int64_t p0 = 0, int64_t p1 = -1)
        : td_(std::move(td)), factors_(numberOfFactors), paths_(numberOfPaths), seed_(seed), seedMode_(seedMode), factory_(std::move(factory)),
          p0_(p0), p1_(p1 < 0 ? numberOfPaths : p1) {}
    RV getBrownianIncrement(int timeIndex, int factor) override {
        std::lock_guard<std::mutex> lock(mu_);
        if (inc_.empty()) generate();
        return inc_[(size_t)timeIndex * factors_ + factor];
    }
    const TimeDiscretization& getTimeDiscretization() const override { return td_; }
    int getNumberOfFactors() const override { return factors_; }
    int64_t getNumberOfPaths() const override { return paths_; }
    RV getRandomVariableForConstant(double value) const override { return factory_->createRandomVariable(value); }
private:
    void generate() {
        const int T = td_.getNumberOfTimeSteps();
        std::vector<double> sq((size_t)T);
        for (int t = 0; t < T; t++) sq[(size_t)t] = std::sqrt(td_.getTimeStep(t));
        const int64_t np = p1_ - p0_;
        std::vector<double> all((size_t)T * factors_ * (size_t)np);
        orc_brownian_f64(seedMode_, seed_, T, factors_, p0_, p1_, sq.data(), all.data());
        inc_.resize((size_t)T * factors_);
        for (int t = 0; t < T; t++)
            for (int f = 0; f < factors_; f++)
                inc_[(size_t)t * factors_ + f] = factory_->createRandomVariable(td_.getTime(t + 1), all.data() + ((size_t)t * factors_ + f) * (size_t)np, np);
    }
    TimeDiscretization td_;
    int factors_; int64_t paths_; int seed_, seedMode_;
    std::shared_ptr<const RandomVariableFactory> factory_;
    int64_t p0_, p1_;
    std::vector<RV> inc_;
    std::mutex mu_;
};

}  // namespace finmath
