// RandomVariable.hpp — C++ mirror of the finmath-lib interfaces the hot path is written against:
//   net.finmath.stochastic.RandomVariable, net.finmath.montecarlo.RandomVariableFactory,
//   net.finmath.montecarlo.BrownianMotion, ConditionalExpectationEstimator.
// Same method names, argument meaning and error behaviour as the methods overridden in
//   /root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:785-1701 and
//   /root/reference/src/main/java/net/finmath/cuda/cpu/montecarlo/RandomVariableFromFloatArray.java:233-1451.
// Objects are immutable; every operation returns a new object (RVC:64-65).
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

namespace finmath {

class RandomVariable;
using RV = std::shared_ptr<const RandomVariable>;

class ConditionalExpectationEstimator {
public:
    virtual ~ConditionalExpectationEstimator() = default;
    virtual RV getConditionalExpectation(const RV& randomVariable) const = 0;
};

struct UnsupportedOperationException : std::logic_error { using std::logic_error::logic_error; };

class RandomVariable : public std::enable_shared_from_this<RandomVariable> {
public:
    virtual ~RandomVariable() = default;
    RV self() const { return shared_from_this(); }

    virtual double getFiltrationTime() const = 0;
    virtual int getTypePriority() const = 0;
    virtual bool isDeterministic() const = 0;
    virtual int64_t size() const = 0;
    virtual double get(int64_t pathOrState) const = 0;
    virtual double doubleValue() const = 0;
    virtual std::vector<double> getRealizations() const = 0;

    virtual double getMin() const = 0;
    virtual double getMax() const = 0;
    virtual double getAverage() const = 0;
    virtual double getAverage(const RV& probabilities) const = 0;
    virtual double getVariance() const = 0;
    virtual double getVariance(const RV& probabilities) const = 0;
    virtual double getSampleVariance() const = 0;
    virtual double getStandardDeviation() const { return isDeterministic() ? 0.0 : std::sqrt(getVariance()); }
    virtual double getStandardError() const { return isDeterministic() ? 0.0 : (size() == 0 ? NAN : getStandardDeviation() / std::sqrt((double)size())); }
    virtual double getQuantile(double quantile) const = 0;
    virtual double getQuantileExpectation(double quantileStart, double quantileEnd) const = 0;

    virtual RV cap(double cap) const = 0;
    virtual RV floor(double floor) const = 0;
    virtual RV add(double value) const = 0;
    virtual RV sub(double value) const = 0;
    virtual RV bus(double value) const = 0;
    virtual RV mult(double value) const = 0;
    virtual RV div(double value) const = 0;
    virtual RV vid(double value) const = 0;
    virtual RV pow(double exponent) const = 0;
    virtual RV average() const = 0;
    virtual RV squared() const = 0;
    virtual RV sqrt() const = 0;
    virtual RV exp() const = 0;
    virtual RV log() const = 0;
    virtual RV sin() const = 0;
    virtual RV cos() const = 0;
    virtual RV invert() const = 0;
    virtual RV abs() const = 0;
    virtual RV isNaN() const = 0;

    virtual RV add(const RV& randomVariable) const = 0;
    virtual RV sub(const RV& randomVariable) const = 0;
    virtual RV bus(const RV& randomVariable) const = 0;
    virtual RV mult(const RV& randomVariable) const = 0;
    virtual RV div(const RV& randomVariable) const = 0;
    virtual RV vid(const RV& randomVariable) const = 0;
    virtual RV cap(const RV& randomVariable) const = 0;
    virtual RV floor(const RV& randomVariable) const = 0;

    virtual RV accrue(const RV& rate, double periodLength) const = 0;
    virtual RV discount(const RV& rate, double periodLength) const = 0;
    virtual RV choose(const RV& valueIfTriggerNonNegative, const RV& valueIfTriggerNegative) const = 0;
    virtual RV addProduct(const RV& factor1, double factor2) const = 0;
    virtual RV addProduct(const RV& factor1, const RV& factor2) const = 0;
    virtual RV addRatio(const RV& numerator, const RV& denominator) const = 0;
    virtual RV subRatio(const RV& numerator, const RV& denominator) const = 0;

    // interface defaults
    virtual RV addSumProduct(const std::vector<RV>& factor1, const std::vector<RV>& factor2) const {   // RVF:1384-1392
        RV result = self();
        for (size_t i = 0; i < factor1.size(); i++) result = result->addProduct(factor1[i], factor2[i]);
        return result;
    }
    virtual RV getConditionalExpectation(const ConditionalExpectationEstimator& estimator) const {      // RVF:860-864
        return estimator.getConditionalExpectation(self());
    }
};

class RandomVariableFactory {
public:
    virtual ~RandomVariableFactory() = default;
    virtual RV createRandomVariable(double time, double value) const = 0;
    virtual RV createRandomVariable(double time, const double* values, int64_t n) const = 0;
    RV createRandomVariable(double value) const { return createRandomVariable(-INFINITY, value); }   // AbstractRandomVariableFactory
    RV createRandomVariable(double time, const std::vector<double>& values) const { return createRandomVariable(time, values.data(), (int64_t)values.size()); }
};

// net.finmath.time.TimeDiscretizationFromArray (the part the drivers use)
class TimeDiscretization {
public:
    TimeDiscretization() = default;
    TimeDiscretization(double initial, int numberOfTimeSteps, double deltaT) {
        for (int i = 0; i <= numberOfTimeSteps; i++) times_.push_back(initial + i * deltaT);
    }
    explicit TimeDiscretization(std::vector<double> times) : times_(std::move(times)) {}
    int getNumberOfTimeSteps() const { return (int)times_.size() - 1; }
    int getNumberOfTimes() const { return (int)times_.size(); }
    double getTime(int i) const { return times_[i]; }
    double getTimeStep(int i) const { return times_[i + 1] - times_[i]; }
    // java.util.Arrays.binarySearch convention: -(insertion point) - 1 when the time is not a grid point
    int getTimeIndex(double time) const {
        int lo = 0, hi = (int)times_.size() - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) / 2;
            if (std::fabs(times_[mid] - time) < 1e-12) return mid;
            if (times_[mid] < time) lo = mid + 1; else hi = mid - 1;
        }
        return -lo - 1;
    }
private:
    std::vector<double> times_;
};

// net.finmath.montecarlo.BrownianMotion
class BrownianMotion {
public:
    virtual ~BrownianMotion() = default;
    virtual RV getBrownianIncrement(int timeIndex, int factor) = 0;
    RV getIncrement(int timeIndex, int factor) { return getBrownianIncrement(timeIndex, factor); }
    virtual const TimeDiscretization& getTimeDiscretization() const = 0;
    virtual int getNumberOfFactors() const = 0;
    virtual int64_t getNumberOfPaths() const = 0;
    virtual RV getRandomVariableForConstant(double value) const = 0;
};

}  // namespace finmath
