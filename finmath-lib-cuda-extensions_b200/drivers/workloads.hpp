// workloads.hpp — stand-ins for the finmath-lib callers of the hot path, written ONLY against the RandomVariable /
// RandomVariableFactory / BrownianMotion interfaces (include/finmath/RandomVariable.hpp), so that the same source runs
// on RandomVariableCuda (the product) and on the CPU oracle twin. finmath-lib 5.1.3 itself is an un-vendored Maven
// dependency of the reference (pom.xml:72-76) and there is no JVM here, so the op sequences of its Euler scheme,
// LIBOR market model, swaption and Bermudan products are restated from the library's published source
// (SURVEY.md Appendix C); the reference call sites are cited per function.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <exception>
#include <memory>
#include <vector>

#include "../../include/finmath/RandomVariable.hpp"

namespace workloads {

using namespace finmath;
using FactoryPtr = std::shared_ptr<const RandomVariableFactory>;
using RegressionFn = std::function<void(const std::vector<RV>& basis, const RV& y, std::vector<double>& XtX, std::vector<double>& XtY)>;

// ---------------------------------------------------------------------------------------------------------------------
// Black-Scholes European call by Euler Monte-Carlo — MonteCarloBlackScholesModelTest.java:62-76, 126-156
//   EulerSchemeFromProcessModel(BlackScholesModel(S0, r, sigma), brownian): log state space,
//   X' = X.addProduct(drift, dt).addSumProduct([sigma], [dW]),  S = exp(X);  numeraire N(t) = exp(r t)
//   value = asset.sub(K).floor(0).div(N(T)).mult(N(0)).getAverage()
// ---------------------------------------------------------------------------------------------------------------------
struct BlackScholesResult { double value; double analytic; };

inline double normal_cdf(double x) { return 0.5 * std::erfc(-x / std::sqrt(2.0)); }

inline BlackScholesResult black_scholes_call(const FactoryPtr& factory, BrownianMotion& brownian, double S0, double r, double sigma,
                                             double maturity, double strike) {
    const TimeDiscretization& td = brownian.getTimeDiscretization();
    const int maturityIndex = td.getTimeIndex(maturity);
    RV state = factory->createRandomVariable(0.0, std::log(S0));                 // BlackScholesModel.getInitialState: log(S0)
    const RV drift = factory->createRandomVariable(r - 0.5 * sigma * sigma);     // BlackScholesModel.getDrift
    const std::vector<RV> factorLoading{factory->createRandomVariable(sigma)};   // BlackScholesModel.getFactorLoading
    RV asset = state->exp();
    for (int t = 1; t <= maturityIndex; t++) {                                   // EulerSchemeFromProcessModel.doPrecalculateProcess
        const double dt = td.getTimeStep(t - 1);
        state = state->addProduct(drift, dt);
        state = state->addSumProduct(factorLoading, {brownian.getBrownianIncrement(t - 1, 0)});
        asset = state->exp();                                                    // applyStateSpaceTransform
    }
    const RV numeraireAtPayment = factory->createRandomVariable(maturity, std::exp(r * maturity));   // MonteCarloBlackScholesModelTest.java:140
    const RV numeraireAtEval = factory->createRandomVariable(0.0, 1.0);                              // :141
    const RV payoff = asset->sub(strike)->floor(0.0);                                                // :143
    BlackScholesResult res;
    res.value = payoff->div(numeraireAtPayment)->mult(numeraireAtEval)->getAverage();                // :144
    const double dp = (std::log(S0 / strike) + (r + 0.5 * sigma * sigma) * maturity) / (sigma * std::sqrt(maturity));
    const double dm = dp - sigma * std::sqrt(maturity);
    res.analytic = S0 * normal_cdf(dp) - strike * std::exp(-r * maturity) * normal_cdf(dm);          // AnalyticFormulas.blackScholesOptionValue
    return res;
}

// ---------------------------------------------------------------------------------------------------------------------
// LIBOR market model, SPOT measure, NORMAL state space, piecewise-constant volatility, one-factor exponential-decay
// correlation — the model of LIBORMarketModelCalibrationATMTest.java:275-314.
// ---------------------------------------------------------------------------------------------------------------------
struct PiecewiseConstantVolatility {
    // LIBORVolatilityModelPiecewiseConstant(timeDisc, liborDisc, simulationTimeGrid, timeToMaturityGrid, value) (T-ATM:287)
    std::vector<double> simGrid{0.0, 1.0, 2.0, 5.0, 10.0, 20.0, 30.0, 40.0}, ttmGrid{0.0, 1.0, 2.0, 5.0, 10.0, 20.0, 30.0, 40.0};
    std::vector<std::vector<int>> index;     // [simBucket][ttmBucket] -> parameter index or -1
    std::vector<double> param;

    explicit PiecewiseConstantVolatility(double value = 0.005) {
        const double maxMaturity = ttmGrid.back();
        int k = 0;
        index.assign(simGrid.size(), std::vector<int>(ttmGrid.size(), -1));
        for (size_t s = 0; s < simGrid.size(); s++)
            for (size_t m = 0; m < ttmGrid.size(); m++)
                if (simGrid[s] + ttmGrid[m] <= maxMaturity) index[s][m] = k++;
        param.assign((size_t)k, value);
    }
    TimeDiscretization simDisc{simGrid}, ttmDisc{ttmGrid};    // the model's two TimeDiscretization members
    static int bucket(const TimeDiscretization& td, double x) {
        // TimeDiscretization.getTimeIndex with the "-idx-1-1" fix-up of LIBORVolatilityModelPiecewiseConstant.getVolatility
        int i = td.getTimeIndex(x);
        if (i < 0) i = -i - 1 - 1;
        if (i < 0) i = 0;
        if (i >= td.getNumberOfTimes()) i = td.getNumberOfTimes() - 1;
        return i;
    }
    double volatility(double time, double maturity) const {
        const double ttm = maturity - time;
        if (ttm <= 0) return 0.0;
        int s = bucket(simDisc, time), m = bucket(ttmDisc, ttm);
        while (m > 0 && index[(size_t)s][(size_t)m] < 0) m--;
        return param[(size_t)index[(size_t)s][(size_t)m]];
    }
};

class LIBORMarketModel {
public:
    LIBORMarketModel(FactoryPtr factory, std::shared_ptr<BrownianMotion> brownian, std::vector<double> forwardRates,
                     PiecewiseConstantVolatility vol)
        : factory_(std::move(factory)), brownian_(std::move(brownian)), L0_(std::move(forwardRates)), vol_(std::move(vol)),
          td_(brownian_->getTimeDiscretization()) {}

    int getNumberOfLibors() const { return (int)L0_.size(); }
    const TimeDiscretization& getTimeDiscretization() const { return td_; }
    double periodLength(int i) const { return td_.getTimeStep(i); }        // liborPeriodDiscretization == timeDiscretization (T-ATM:278)
    const std::vector<double>& forwardRates() const { return L0_; }
    int64_t numberOfPaths() const { return brownian_->getNumberOfPaths(); }
    const FactoryPtr& factory() const { return factory_; }

    // EulerSchemeFromProcessModel.doPrecalculateProcess (finmath-lib), every arithmetic step one RandomVariable call
    void simulate() {
        const int T = td_.getNumberOfTimeSteps(), NC = getNumberOfLibors(), F = brownian_->getNumberOfFactors();
        process_.assign((size_t)T + 1, std::vector<RV>((size_t)NC));
        numeraire_.clear();
        for (int i = 0; i < NC; i++) process_[0][(size_t)i] = factory_->createRandomVariable(0.0, L0_[(size_t)i]);   // getInitialState
        std::vector<RV> dW((size_t)F);
        for (int t = 1; t <= T; t++) {
            const double dt = td_.getTimeStep(t - 1);
            const std::vector<RV> drift = getDrift(t - 1, process_[(size_t)t - 1]);
            for (int f = 0; f < F; f++) dW[(size_t)f] = brownian_->getBrownianIncrement(t - 1, f);
            for (int i = 0; i < NC; i++) {
                if (!drift[(size_t)i]) { process_[(size_t)t][(size_t)i] = process_[(size_t)t - 1][(size_t)i]; continue; }
                const std::vector<RV> fl = getFactorLoading(t - 1, i);
                RV state = process_[(size_t)t - 1][(size_t)i]->addProduct(drift[(size_t)i], dt);   // "mu DeltaT"
                state = state->addSumProduct(fl, dW);                                             // diffusion
                process_[(size_t)t][(size_t)i] = state;                                           // NORMAL state space: identity transform
            }
        }
    }

    RV getLIBOR(int timeIndex, int liborIndex) const { return process_[(size_t)timeIndex][(size_t)liborIndex]; }

    // LIBORMarketModelFromCovarianceModel.getNumeraire, SPOT measure: N(T_k) = N(T_{k-1}).accrue(L_{k-1}(T_{k-1}), delta)
    // (cached per model like the library's ConcurrentHashMap of numeraires: valuation threads share it)
    RV getNumeraire(int liborIndex) {
        std::lock_guard<std::recursive_mutex> lock(numeraire_mu_);
        auto it = numeraire_.find(liborIndex);
        if (it != numeraire_.end()) return it->second;
        RV n;
        if (liborIndex == 0) n = factory_->createRandomVariable(0.0, 1.0);
        else n = getNumeraire(liborIndex - 1)->accrue(getLIBOR(liborIndex - 1, liborIndex - 1), periodLength(liborIndex - 1));
        numeraire_[liborIndex] = n;
        return n;
    }

    PiecewiseConstantVolatility& volatilityModel() { return vol_; }

private:
    // LIBORCovarianceModelFromVolatilityAndCorrelation.getFactorLoading: vol_i(t) * factorMatrix[i][f]; with ONE factor the
    // reduced, renormalised exponential-decay correlation (T-ATM:288) has factorMatrix[i][0] = 1.
    std::vector<RV> getFactorLoading(int timeIndex, int component) const {
        const double v = vol_.volatility(td_.getTime(timeIndex), td_.getTime(component));
        std::vector<RV> fl;
        for (int f = 0; f < brownian_->getNumberOfFactors(); f++) fl.push_back(factory_->createRandomVariable(f == 0 ? v : 0.0));
        return fl;
    }

    // LIBORMarketModelFromCovarianceModel.getDrift, Measure.SPOT, StateSpace.NORMAL
    std::vector<RV> getDrift(int timeIndex, const std::vector<RV>& realization) const {
        const int NC = getNumberOfLibors(), F = brownian_->getNumberOfFactors();
        const int first = timeIndex + 1;                       // getLiborPeriodIndex(time) + 1 on the common grid
        const RV zero = factory_->createRandomVariable(0.0);
        std::vector<RV> drift((size_t)NC);
        for (int i = first; i < NC; i++) drift[(size_t)i] = zero;
        std::vector<RV> covarianceFactorSums((size_t)F, zero);
        for (int i = first; i < NC; i++) {
            const double p = periodLength(i);
            const RV forwardRate = realization[(size_t)i];
            const RV oneStepMeasureTransform = factory_->createRandomVariable(p)->discount(forwardRate, p);
            const std::vector<RV> fl = getFactorLoading(timeIndex, i);
            for (int f = 0; f < F; f++) {
                covarianceFactorSums[(size_t)f] = covarianceFactorSums[(size_t)f]->add(oneStepMeasureTransform->mult(fl[(size_t)f]));
                drift[(size_t)i] = drift[(size_t)i]->addProduct(covarianceFactorSums[(size_t)f], fl[(size_t)f]);
            }
        }
        return drift;
    }

    FactoryPtr factory_;
    std::shared_ptr<BrownianMotion> brownian_;
    std::vector<double> L0_;
    PiecewiseConstantVolatility vol_;
    TimeDiscretization td_;
    std::vector<std::vector<RV>> process_;
    std::map<int, RV> numeraire_;
    std::recursive_mutex numeraire_mu_;
};

// analytic curve quantities from the initial forward rates (ForwardCurve / DiscountCurveFromForwardCurve, T-ATM:353-355)
struct Curve {
    std::vector<double> discount;      // P(0, T_i), i = 0..NC
    explicit Curve(const std::vector<double>& L0, double delta) {
        discount.assign(L0.size() + 1, 1.0);
        for (size_t i = 0; i < L0.size(); i++) discount[i + 1] = discount[i] / (1.0 + L0[i] * delta);
    }
    double annuity(int e, int m, double delta) const { double a = 0; for (int j = 0; j < m; j++) a += delta * discount[(size_t)(e + j + 1)]; return a; }
    double parSwapRate(int e, int m, double delta) const { return (discount[(size_t)e] - discount[(size_t)(e + m)]) / annuity(e, m, delta); }
};

struct SwaptionSpec { int exerciseIndex; int numberOfPeriods; double strike; double targetVolatility; };

// net.finmath.montecarlo.interestrate.products.Swaption.getValue (used by SwaptionSimple, T-ATM createCalibrationItem)
inline RV swaption_values(LIBORMarketModel& model, const SwaptionSpec& s) {
    const int e = s.exerciseIndex;
    RV valueOfSwapAtExerciseDate = model.factory()->createRandomVariable(0.0);
    for (int period = s.numberOfPeriods - 1; period >= 0; period--) {
        const int i = e + period;
        const double periodLength = model.periodLength(i);
        const RV libor = model.getLIBOR(e, i);
        const RV payoff = libor->sub(s.strike)->mult(periodLength);
        valueOfSwapAtExerciseDate = valueOfSwapAtExerciseDate->add(payoff);
        valueOfSwapAtExerciseDate = valueOfSwapAtExerciseDate->discount(libor, periodLength);
    }
    RV values = valueOfSwapAtExerciseDate->floor(0.0);
    const double w = 1.0 / (double)model.numberOfPaths();
    const RV numeraire = model.getNumeraire(e);
    const RV monteCarloProbabilities = model.factory()->createRandomVariable(w);            // getMonteCarloWeights
    values = values->div(numeraire)->mult(monteCarloProbabilities);
    const RV numeraireAtZero = model.getNumeraire(0);
    const RV monteCarloProbabilitiesAtZero = model.factory()->createRandomVariable(w);
    values = values->mult(numeraireAtZero)->div(monteCarloProbabilitiesAtZero);
    return values;
}

// the 154-product ATM calibration set of LIBORMarketModelCalibrationATMTest.java:188-268 on the idealised 0.25y grid:
// expiries {1,2,3,4,5,7,10,15,20,25,30}Y x tenors {1..10,15,20,25,30}Y, swap period 0.5; products whose swap ends after
// the 40y simulation horizon are dropped (their valuation throws in the reference and is swallowed, T-ATM:381-400).
inline std::vector<SwaptionSpec> atm_calibration_products(const std::vector<double>& L0, double delta, int horizonPeriods) {
    static const int expiries[] = {1, 2, 3, 4, 5, 7, 10, 15, 20, 25, 30};
    static const int tenors[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 15, 20, 25, 30};
    static const double vols[11][14] = {   // atmNormalVolatilities rows for expiries 1Y..30Y (T-ATM:219-236)
        {0.00205, 0.00235, 0.00272, 0.0032, 0.00368, 0.00406, 0.00447, 0.00484, 0.00515, 0.00544, 0.00602, 0.00629, 0.0064, 0.00646},
        {0.00279, 0.00319, 0.0036, 0.00396, 0.00436, 0.00469, 0.00503, 0.0053, 0.00557, 0.00582, 0.00616, 0.00628, 0.00638, 0.00641},
        {0.00379, 0.00406, 0.00439, 0.00472, 0.00504, 0.00532, 0.0056, 0.00582, 0.00602, 0.00617, 0.0063, 0.00636, 0.00638, 0.00639},
        {0.00471, 0.00489, 0.00511, 0.00539, 0.00563, 0.00583, 0.006, 0.00618, 0.0063, 0.00644, 0.00641, 0.00638, 0.00635, 0.00634},
        {0.00544, 0.00557, 0.00572, 0.00591, 0.00604, 0.00617, 0.0063, 0.00641, 0.00651, 0.00661, 0.00645, 0.00634, 0.00627, 0.00624},
        {0.00625, 0.00632, 0.00638, 0.00644, 0.0065, 0.00655, 0.00661, 0.00667, 0.00672, 0.00673, 0.00634, 0.00614, 0.00599, 0.00593},
        {0.00664, 0.00671, 0.00675, 0.00676, 0.00676, 0.00675, 0.00676, 0.00674, 0.00672, 0.00669, 0.00616, 0.00586, 0.00569, 0.00558},
        {0.00647, 0.00651, 0.00651, 0.00651, 0.00652, 0.00649, 0.00645, 0.0064, 0.00637, 0.00631, 0.00576, 0.00534, 0.00512, 0.00495},
        {0.00615, 0.0062, 0.00618, 0.00613, 0.0061, 0.00607, 0.00602, 0.00596, 0.00591, 0.00586, 0.00536, 0.00491, 0.00469, 0.0045},
        {0.00578, 0.00583, 0.00579, 0.00574, 0.00567, 0.00562, 0.00556, 0.00549, 0.00545, 0.00538, 0.00493, 0.00453, 0.00435, 0.0042},
        {0.00542, 0.00547, 0.00539, 0.00532, 0.00522, 0.00516, 0.0051, 0.00504, 0.005, 0.00495, 0.00454, 0.00418, 0.00404, 0.00394}};
    Curve curve(L0, delta);
    std::vector<SwaptionSpec> out;
    for (int a = 0; a < 11; a++)
        for (int b = 0; b < 14; b++) {
            const int e = (int)std::lround(expiries[a] / delta), m = (int)std::lround(tenors[b] / delta);
            if (e + m > horizonPeriods) continue;
            out.push_back(SwaptionSpec{e, m, curve.parSwapRate(e, m, delta), vols[a][b]});   // moneyness 0 (T-ATM:259)
        }
    return out;
}

// Bachelier implied volatility of an ATM payer swaption: value = annuity * sigma * sqrt(T / (2 pi))
inline double atm_normal_implied_vol(double value, double annuity, double optionMaturity) {
    return value / (annuity * std::sqrt(optionMaturity / (2.0 * M_PI)));
}

// synthetic upward-sloping forward curve (the reference calibrates its curve from market swap quotes, T-ATM:526-663,
// with finmath's analytic curve solver: out of scope; only the SHAPE of the workload matters for the hot path)
inline std::vector<double> synthetic_forward_rates(int n, double delta) {
    std::vector<double> L((size_t)n);
    for (int i = 0; i < n; i++) { const double t = i * delta; L[(size_t)i] = 0.005 + 0.02 * (1.0 - std::exp(-t / 8.0)); }
    return L;
}

// The EUR curve of LIBORMarketModelCalibrationATMTest.java:526-663: a single discount curve (log-linear interpolation of the
// discount factors, constant extrapolation) calibrated to 21 par swap rates, fixed leg annual against 6M floating, and the 6M
// forward curve derived from it. finmath-lib's schedule generator, business-day calendar and day-count conventions are not
// part of the reference tree: the schedules are idealised here (year fractions = period lengths, no spot lag), the pillars sit
// at the swap maturities, and each pillar is solved in turn (the curve left of a pillar does not depend on it), which is what
// the library's multi-dimensional solver converges to for this triangular system. The hot path only sees the resulting
// initial forward rates.
struct MarketCurveATM {
    std::vector<double> pillarT{0.0}, pillarLogDF{0.0};
    double logDF(double t) const {
        if (t <= 0.0) return 0.0;
        if (t >= pillarT.back()) return pillarLogDF.back();                         // ExtrapolationMethod.CONSTANT (T-ATM:612)
        size_t k = 1;
        while (pillarT[k] < t) k++;
        const double w = (t - pillarT[k - 1]) / (pillarT[k] - pillarT[k - 1]);
        return (1.0 - w) * pillarLogDF[k - 1] + w * pillarLogDF[k];                  // InterpolationEntity.LOG_OF_VALUE, LINEAR (T-ATM:600-613)
    }
    double discountFactor(double t) const { return std::exp(logDF(t)); }
    MarketCurveATM() {
        static const double maturity[] = {0.5, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15, 20, 25, 30, 35, 40, 45, 50};          // T-ATM:527
        static const double rates[] = {-0.00216, -0.00208, -0.00222, -0.00216, -0.0019, -0.0014, -0.00072, 0.00011, 0.00103, 0.00196, 0.00285,
                                       0.00367, 0.0044, 0.00604, 0.00733, 0.00767, 0.00773, 0.00765, 0.00752, 0.007138, 0.007};    // T-ATM:532
        for (int i = 0; i < 21; i++) {
            const double T = maturity[i], S = rates[i];
            pillarT.push_back(T); pillarLogDF.push_back(pillarLogDF.back());
            // par condition of a swap from 0 to T, single curve: 1 - P(T) = S * sum_j tau_j P(t_j), fixed payments yearly (one stub of T < 1)
            auto f = [&](double x) {
                pillarLogDF.back() = x;
                double annuity = 0.0;
                if (T < 1.0) annuity = T * discountFactor(T);
                else for (int j = 1; j <= (int)std::lround(T); j++) annuity += discountFactor((double)j);
                return 1.0 - discountFactor(T) - S * annuity;
            };
            double lo = pillarLogDF[pillarLogDF.size() - 2] - 1.0, hi = pillarLogDF[pillarLogDF.size() - 2] + 1.0;      // f is increasing in -x
            for (int it = 0; it < 200; it++) { const double mid = 0.5 * (lo + hi); if (f(mid) > 0.0) lo = mid; else hi = mid; }
            pillarLogDF.back() = 0.5 * (lo + hi);
        }
    }
    // ForwardCurveFromDiscountCurve(discountCurve, "6M"): L(T_i) = (P(T_i) / P(T_i + delta) - 1) / delta
    std::vector<double> forwardRates(int n, double delta) const {
        std::vector<double> L((size_t)n);
        for (int i = 0; i < n; i++) L[(size_t)i] = (discountFactor(i * delta) / discountFactor((i + 1) * delta) - 1.0) / delta;
        return L;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// net.finmath.optimizer.LevenbergMarquardt as LIBORMarketModelCalibrationATMTest.java:317-340 configures it
// (RegularizationMethod.LEVENBERG, lambda 0.1, <= 200 iterations, accuracy 1e-7 on the change of the root mean squared error,
// ONE thread, finite-difference derivatives with parameterStep 1e-4): restated from the library's published algorithm
// (finmath-lib 5.1.3 is an un-vendored dependency). One iteration = one evaluation of the objective at the trial point and,
// after an accepted step, #parameters more for the forward-difference Jacobian — every evaluation is one full Monte-Carlo
// simulation + valuation of all calibration products: the hot path.
// ---------------------------------------------------------------------------------------------------------------------
struct LevenbergMarquardtResult {
    std::vector<double> parameters, values;
    int iterations = 0, evaluations = 0;
    double rootMeanSquaredError = 0.0;
};

inline LevenbergMarquardtResult levenberg_marquardt(const std::function<std::vector<double>(const std::vector<double>&)>& objective,
                                                    std::vector<double> initialParameters, const std::vector<double>& targetValues,
                                                    int maxIteration, double errorTolerance, double lambda, double parameterStep) {
    const size_t np = initialParameters.size(), nv = targetValues.size();
    const double lambdaDivisor = 1.3, lambdaMultiplicator = 2.0;
    LevenbergMarquardtResult res;
    auto meanSquaredError = [&](const std::vector<double>& v) {
        double e = 0.0;
        for (size_t k = 0; k < nv; k++) { const double d = v[k] - targetValues[k]; e += d * d; }      // weights 1 (T-ATM:262)
        return e / (double)nv;
    };
    std::vector<double> parameterCurrent = initialParameters, parameterTest = initialParameters, valueCurrent(nv, NAN), valueTest;
    std::vector<std::vector<double>> derivativeCurrent(np, std::vector<double>(nv, 0.0));
    double errorMeanSquaredCurrent = INFINITY, errorRootMeanSquaredChange = INFINITY;
    bool derivativeValid = false;
    int iteration = 0;
    for (;;) {
        iteration++;
        valueTest = objective(parameterTest); res.evaluations++;
        const double errorMeanSquaredTest = meanSquaredError(valueTest);
        if (errorMeanSquaredTest < errorMeanSquaredCurrent) {              // NaN compares false: a rejected point
            errorRootMeanSquaredChange = std::sqrt(errorMeanSquaredCurrent) - std::sqrt(errorMeanSquaredTest);
            parameterCurrent = parameterTest; valueCurrent = valueTest; errorMeanSquaredCurrent = errorMeanSquaredTest;
            derivativeValid = false;
            lambda /= lambdaDivisor;
        } else {
            errorRootMeanSquaredChange = std::sqrt(errorMeanSquaredTest) - std::sqrt(errorMeanSquaredCurrent);
            lambda *= lambdaMultiplicator;
        }
        if (iteration > maxIteration || errorRootMeanSquaredChange <= errorTolerance) break;
        if (!derivativeValid) {                                            // forward differences, one simulation per parameter
            for (size_t i = 0; i < np; i++) {
                std::vector<double> p = parameterCurrent;
                p[i] += parameterStep;
                const std::vector<double> v = objective(p); res.evaluations++;
                for (size_t k = 0; k < nv; k++) derivativeCurrent[i][k] = (v[k] - valueCurrent[k]) / parameterStep;
            }
            derivativeValid = true;
        }
        // (J^T J + lambda I) delta = J^T (target - value), solved by Gaussian elimination with partial pivoting
        std::vector<double> H(np * np), beta(np);
        for (size_t i = 0; i < np; i++) {
            double b = 0.0;
            for (size_t k = 0; k < nv; k++) b += (targetValues[k] - valueCurrent[k]) * derivativeCurrent[i][k];
            beta[i] = b;
            for (size_t j = 0; j <= i; j++) {
                double a = 0.0;
                for (size_t k = 0; k < nv; k++) a += derivativeCurrent[i][k] * derivativeCurrent[j][k];
                if (i == j) a += lambda;                                   // RegularizationMethod.LEVENBERG
                H[i * np + j] = a; H[j * np + i] = a;
            }
        }
        std::vector<double> delta = beta;
        for (size_t c = 0; c < np; c++) {
            size_t piv = c;
            for (size_t r = c + 1; r < np; r++) if (std::fabs(H[r * np + c]) > std::fabs(H[piv * np + c])) piv = r;
            if (piv != c) { for (size_t j = 0; j < np; j++) std::swap(H[c * np + j], H[piv * np + j]); std::swap(delta[c], delta[piv]); }
            const double d = H[c * np + c];
            if (d == 0.0) continue;
            for (size_t r = c + 1; r < np; r++) {
                const double m = H[r * np + c] / d;
                if (m == 0.0) continue;
                for (size_t j = c; j < np; j++) H[r * np + j] -= m * H[c * np + j];
                delta[r] -= m * delta[c];
            }
        }
        for (size_t c = np; c-- > 0;) {
            double x = delta[c];
            for (size_t j = c + 1; j < np; j++) x -= H[c * np + j] * delta[j];
            delta[c] = H[c * np + c] != 0.0 ? x / H[c * np + c] : 0.0;
        }
        for (size_t i = 0; i < np; i++) parameterTest[i] = parameterCurrent[i] + delta[i];
    }
    res.parameters = parameterCurrent; res.values = valueCurrent; res.iterations = iteration;
    res.rootMeanSquaredError = std::sqrt(errorMeanSquaredCurrent);
    return res;
}

// one pass of the calibration inner loop: simulate the model, value every calibration product.
// threads > 1: the products are valued by that many host threads (product k by thread k mod threads), the way the
// library's calibration does with numberOfThreads > 1; the reference test runs with ONE thread (T-ATM:319), the default.
// price_products: how the averages are taken.
//   false (default, T-ATM:261,511): products in ValueUnit.VOLATILITYNORMAL — SwaptionSimple.getValue() takes value.getAverage()
//         itself to imply the volatility, i.e. one blocking reduction per product, interleaved with the recording of the next;
//   true: products in ValueUnit.VALUE (the alternative LIBORMarketModelCalibrationTest.java:151-154 names) — the objective
//         function of AbstractLIBORCovarianceModelParametric.getCloneCalibrated first collects all product value vectors and
//         then calls getAverage() on each in a second loop (finmath-lib 5.1.3 is not in /root/reference: restated from
//         memory). The runtime then sums all of them in a few launches (fmcuda: batched averages).
inline std::vector<double> lmm_value_products(LIBORMarketModel& model, const std::vector<SwaptionSpec>& products, int threads = 1, bool price_products = false) {
    model.simulate();
    std::vector<double> values(products.size());
    if (threads <= 1) {
        if (price_products) {
            std::vector<RV> v(products.size());
            for (size_t k = 0; k < products.size(); k++) v[k] = swaption_values(model, products[k]);
            for (size_t k = 0; k < products.size(); k++) values[k] = v[k]->getAverage();
            return values;
        }
        for (size_t k = 0; k < products.size(); k++) values[k] = swaption_values(model, products[k])->getAverage();
        return values;
    }
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> errors((size_t)threads);
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t] {
            try {
                for (size_t k = (size_t)t; k < products.size(); k += (size_t)threads) values[k] = swaption_values(model, products[k])->getAverage();
            } catch (...) { errors[(size_t)t] = std::current_exception(); }
        });
    for (auto& th : pool) th.join();
    for (auto& e : errors) if (e) std::rethrow_exception(e);
    return values;
}

// ---------------------------------------------------------------------------------------------------------------------
// Bermudan swaption by backward induction with regression — net.finmath.montecarlo.interestrate.products.BermudanSwaption
// (config 3 of BASELINE.json; needs choose(), which the reference GPU class lacks, RVC:1632-1635)
// ---------------------------------------------------------------------------------------------------------------------
class RegressionEstimator : public ConditionalExpectationEstimator {
public:
    RegressionEstimator(std::vector<RV> basis, RegressionFn normalEquations) : basis_(std::move(basis)), neq_(std::move(normalEquations)) {}
    std::vector<double> getLinearRegressionParameters(const RV& y) const {
        std::vector<double> XtX, XtY;
        neq_(basis_, y, XtX, XtY);
        return solve_spd(XtX, XtY, (int)basis_.size());
    }
    RV getConditionalExpectation(const RV& y) const override {                  // MonteCarloConditionalExpectationRegression
        const std::vector<double> c = getLinearRegressionParameters(y);
        RV est = basis_[0]->mult(c[0]);
        for (size_t i = 1; i < basis_.size(); i++) est = est->addProduct(basis_[i], c[i]);
        return est;
    }
    // Minimum-norm least-squares solution of the symmetric k x k system (k <= 12, host double), the job finmath-lib gives to
    // commons-math3's SingularValueDecomposition solver: cyclic Jacobi eigen-decomposition A = V diag(w) V^T, components
    // whose eigenvalue is below 1e-10 of the largest are dropped. The basis of a Bermudan regression (1, discount factor,
    // its square, ...) is nearly collinear: the normal equations have condition numbers of 1e12 and beyond, and a plain
    // Cholesky solve turns the last bits of the float products into O(1) coefficient noise (exercise decisions flip).
    static std::vector<double> solve_spd(std::vector<double> A, std::vector<double> b, int k) {
        std::vector<double> V((size_t)k * k, 0.0);
        for (int i = 0; i < k; i++) V[(size_t)i * k + i] = 1.0;
        auto a = [&](int i, int j) -> double& { return A[(size_t)i * k + j]; };
        for (int sweep = 0; sweep < 60; sweep++) {
            double off = 0.0, diag = 0.0;
            for (int i = 0; i < k; i++) { diag += a(i, i) * a(i, i); for (int j = i + 1; j < k; j++) off += a(i, j) * a(i, j); }
            if (off <= 1e-32 * diag) break;
            for (int p = 0; p < k; p++)
                for (int q = p + 1; q < k; q++) {
                    if (a(p, q) == 0.0) continue;
                    const double theta = (a(q, q) - a(p, p)) / (2.0 * a(p, q));
                    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                    const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                    for (int m = 0; m < k; m++) {                      // A <- A J  (columns p, q)
                        const double amp = a(m, p), amq = a(m, q);
                        a(m, p) = c * amp - sn * amq; a(m, q) = sn * amp + c * amq;
                    }
                    for (int m = 0; m < k; m++) {                      // A <- J^T A  (rows p, q)
                        const double apm = a(p, m), aqm = a(q, m);
                        a(p, m) = c * apm - sn * aqm; a(q, m) = sn * apm + c * aqm;
                    }
                    for (int m = 0; m < k; m++) {                      // V <- V J
                        const double vmp = V[(size_t)m * k + p], vmq = V[(size_t)m * k + q];
                        V[(size_t)m * k + p] = c * vmp - sn * vmq; V[(size_t)m * k + q] = sn * vmp + c * vmq;
                    }
                }
        }
        double wmax = 0.0;
        for (int i = 0; i < k; i++) wmax = std::max(wmax, std::fabs(a(i, i)));
        std::vector<double> x((size_t)k, 0.0);
        for (int e = 0; e < k; e++) {
            const double w = a(e, e);
            if (!(std::fabs(w) > 1e-10 * wmax)) continue;
            double proj = 0.0;
            for (int m = 0; m < k; m++) proj += V[(size_t)m * k + e] * b[(size_t)m];
            for (int m = 0; m < k; m++) x[(size_t)m] += V[(size_t)m * k + e] * proj / w;
        }
        return x;
    }
private:
    std::vector<RV> basis_;
    RegressionFn neq_;
};

struct BermudanSpec { int firstExerciseIndex; int lastExerciseIndex; int exerciseStride; int swapEndIndex; double strike; };

// LIBORMonteCarloSimulationFromLIBORModel.getLIBOR(t, periodStart, periodEnd) over several model periods
inline RV libor_over_periods(LIBORMarketModel& model, int timeIndex, int startIndex, int endIndex) {
    RV accrualAccount;
    double length = 0.0;
    for (int j = startIndex; j < endIndex; j++) {
        const double sub = model.periodLength(j);
        const RV l = model.getLIBOR(timeIndex, j);
        accrualAccount = accrualAccount ? accrualAccount->accrue(l, sub) : l->mult(sub)->add(1.0);
        length += sub;
    }
    return accrualAccount->sub(1.0)->div(length);
}

// BermudanSwaption.getValue (finmath-lib): backward over the swap periods; at an exercise date the continuation value is
// compared with the value of the remaining swap through a regression estimate of (continuation - exercise), and
// values = trigger.choose(values, valuesUnderlying). Basis (getRegressionBasisFunctions): 1, short discount factor and
// its square, discount factor to the swap end and its square, 1/numeraire.
inline double bermudan_swaption_value(LIBORMarketModel& model, const BermudanSpec& spec, const RegressionFn& neq) {
    const FactoryPtr& fac = model.factory();
    const double w = 1.0 / (double)model.numberOfPaths();
    RV values = fac->createRandomVariable(0.0);
    RV valuesUnderlying = fac->createRandomVariable(0.0);
    for (int period = spec.swapEndIndex - 1; period >= spec.firstExerciseIndex; period--) {
        const double periodLength = model.periodLength(period);
        const RV libor = model.getLIBOR(period, period);                       // rate at simulation time = fixing date
        RV payoff = libor->sub(spec.strike)->mult(periodLength)->mult(1.0 /* notional */);
        const RV numeraire = model.getNumeraire(period + 1);                   // payment date
        const RV monteCarloProbabilities = fac->createRandomVariable(w);
        payoff = payoff->div(numeraire)->mult(monteCarloProbabilities);
        valuesUnderlying = valuesUnderlying->add(payoff);
        const bool isExercise = period <= spec.lastExerciseIndex && (period - spec.firstExerciseIndex) % spec.exerciseStride == 0;
        if (!isExercise) continue;
        const RV triggerValuesDiscounted = values->sub(valuesUnderlying);
        std::vector<RV> basis;
        basis.push_back(fac->createRandomVariable(model.getTimeDiscretization().getTime(period), 1.0));
        const RV rateShort = model.getLIBOR(period, period);
        const RV discountShort = rateShort->mult(periodLength)->add(1.0)->invert();
        basis.push_back(discountShort);
        basis.push_back(discountShort->pow(2.0));
        double lengthLong = 0.0; for (int j = period; j < spec.swapEndIndex; j++) lengthLong += model.periodLength(j);
        const RV rateLong = libor_over_periods(model, period, period, spec.swapEndIndex);
        const RV discountLong = rateLong->mult(lengthLong)->add(1.0)->invert();
        basis.push_back(discountLong);
        basis.push_back(discountLong->pow(2.0));
        basis.push_back(model.getNumeraire(period)->invert());
        RegressionEstimator estimator(basis, neq);
        const RV triggerValues = triggerValuesDiscounted->getConditionalExpectation(estimator);
        values = triggerValues->choose(values, valuesUnderlying);
    }
    const RV numeraireAtZero = model.getNumeraire(0);
    const RV monteCarloProbabilitiesAtZero = fac->createRandomVariable(w);
    return values->mult(numeraireAtZero)->div(monteCarloProbabilitiesAtZero)->getAverage();
}

}  // namespace workloads
