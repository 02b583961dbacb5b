// RandomVariableFromDoubleArray.hpp — CPU twin of finmath-lib's DEFAULT vector type in C++ (test infrastructure / CPU
// baseline only, NOT product code).
//
// net.finmath.montecarlo.RandomVariableFromDoubleArray (finmath-lib 5.1.3; imported by RandomVariableCuda.java:50 and
// RandomVariableFromFloatArray.java:21, source not vendored in the reference tree) is the class the float variant was
// derived from: the same element-wise loops, every value and every intermediate a Java double, reductions by the same
// Kahan loops (the float class differs by its (float) casts only, cf. RandomVariableFromFloatArray.java:314-420,
// 751-1451). It is what a plain finmath-lib user runs on the CPU (RandomVariableFromArrayFactory), so the north star
// asks for its timing next to the float twin. PARITY UNPINNED: restated from the published source, no fixture for it.
#pragma once
#include <cmath>
#include <vector>

#include "../include/finmath/RandomVariableImpl.hpp"

namespace finmath {

struct DoubleArrayBackend {
    static constexpr int kTypePriority = 1;
    using Vec = std::vector<double>;
    static void release(Vec&) {}
    static Vec from_f64(const double* p, int64_t n) { return Vec(p, p + n); }
    static double un(int op, double a, double s) {
        switch (op) {
        case OP_CAP: return detail::java_min(a, s);
        case OP_FLOOR: return detail::java_max(a, s);
        case OP_ADD: return a + s;
        case OP_SUB: return a - s;
        case OP_BUS: return s - a;
        case OP_MULT: return a * s;
        case OP_DIV: return a / s;
        case OP_VID: return s / a;
        case OP_POW: return std::pow(a, s);
        case OP_SQUARED: return a * a;
        case OP_SQRT: return std::sqrt(a);
        case OP_EXP: return std::exp(a);
        case OP_LOG: return std::log(a);
        case OP_SIN: return std::sin(a);
        case OP_COS: return std::cos(a);
        case OP_INVERT: return 1.0 / a;
        case OP_ABS: return std::fabs(a);
        case OP_ISNAN: return a != a ? 1.0 : 0.0;
        default: throw std::invalid_argument("bad opcode");
        }
    }
    static Vec vs(int op, const Vec& a, double s, int64_t n) { Vec r((size_t)n); for (int64_t i = 0; i < n; i++) r[(size_t)i] = un(op, a[(size_t)i], s); return r; }
    static Vec v(int op, const Vec& a, int64_t n) { return vs(op, a, 0.0, n); }
    static Vec vv(int op, const Vec& a, const Vec& b, int64_t n, int64_t nb) {
        if (n != nb) throw std::out_of_range("operand sizes differ");
        Vec r((size_t)n);
        for (int64_t i = 0; i < n; i++) r[(size_t)i] = un(op, a[(size_t)i], b[(size_t)i]);
        return r;
    }
    static Vec vvs(int op, const Vec& a, const Vec& b, double s, int64_t n) {
        Vec r((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            const double x = a[(size_t)i], y = b[(size_t)i];
            r[(size_t)i] = op == OP_ACCRUE ? x * (1.0 + y * s) : op == OP_DISCOUNT ? x / (1.0 + y * s) : x + y * s;   // addProduct(v, scalar)
        }
        return r;
    }
    static Vec vvv(int op, const Vec& a, const Vec& b, const Vec& c, int64_t n) {
        Vec r((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            const double x = a[(size_t)i], y = b[(size_t)i], z = c[(size_t)i];
            r[(size_t)i] = op == OP_ADDPRODUCT ? x + y * z : op == OP_ADDRATIO ? x + y / z : op == OP_SUBRATIO ? x - y / z : (x >= 0.0 ? y : z);
        }
        return r;
    }
    static Vec choose(const Vec& t, const Vec* a, double sa, const Vec* b, double sb, int64_t n) {
        Vec r((size_t)n);
        for (int64_t i = 0; i < n; i++) r[(size_t)i] = (t[(size_t)i] >= 0.0) ? (a ? (*a)[(size_t)i] : sa) : (b ? (*b)[(size_t)i] : sb);
        return r;
    }
    template <typename F> static double kahan(int64_t n, F term) {
        double sum = 0.0, error = 0.0;
        for (int64_t i = 0; i < n; i++) { const double value = term(i) - error, newSum = sum + value; error = (newSum - sum) - value; sum = newSum; }
        return sum;
    }
    static double reduce(int kind, const Vec& a, int64_t n, const Vec* w) {
        if (n == 0) return kind == R_MIN ? 1.7976931348623157e308 : kind == R_MAX ? -1.7976931348623157e308 : NAN;
        switch (kind) {
        case R_SUM: return kahan(n, [&](int64_t i) { return a[(size_t)i]; });
        case R_AVERAGE: return kahan(n, [&](int64_t i) { return a[(size_t)i]; }) / (double)n;
        case R_VARIANCE: case R_SAMPLE_VARIANCE: {
            const double m = kahan(n, [&](int64_t i) { return a[(size_t)i]; }) / (double)n;
            const double s = kahan(n, [&](int64_t i) { const double d = a[(size_t)i] - m; return d * d; });
            return kind == R_VARIANCE ? s / (double)n : s / (double)(n - 1);
        }
        case R_MIN: { double m = a[0]; for (int64_t i = 1; i < n; i++) m = detail::java_min(m, a[(size_t)i]); return m; }
        case R_MAX: { double m = a[0]; for (int64_t i = 1; i < n; i++) m = detail::java_max(m, a[(size_t)i]); return m; }
        case R_AVERAGE_W: return kahan(n, [&](int64_t i) { return a[(size_t)i] * (*w)[(size_t)i]; });
        case R_VARIANCE_W: {
            const double m = kahan(n, [&](int64_t i) { return a[(size_t)i] * (*w)[(size_t)i]; });
            return kahan(n, [&](int64_t i) { const double d = a[(size_t)i] - m; return d * d * (*w)[(size_t)i]; });
        }
        default: throw std::invalid_argument("bad reduction kind");
        }
    }
    static double quantile(const Vec&, int64_t, double) { throw std::logic_error("not needed by the baseline drivers"); }
    static double quantile_expectation(const Vec&, int64_t, double, double) { throw std::logic_error("not needed by the baseline drivers"); }
    static double get(const Vec& a, int64_t n, int64_t i) { if (i < 0 || i >= n) throw std::out_of_range("index"); return a[(size_t)i]; }
    static std::vector<double> to_f64(const Vec& a, int64_t) { return a; }
};

using RandomVariableFromDoubleArray = RandomVariableImpl<DoubleArrayBackend>;

// net.finmath.montecarlo.RandomVariableFromArrayFactory: finmath-lib's default factory
class RandomVariableFromArrayFactory : public RandomVariableFactory {
public:
    using RandomVariableFactory::createRandomVariable;
    RV createRandomVariable(double time, double value) const override { return RandomVariableFromDoubleArray::of(time, value); }
    RV createRandomVariable(double time, const double* values, int64_t n) const override { return RandomVariableFromDoubleArray::of(time, values, n); }
};

// MonteCarloConditionalExpectationRegression on doubles: XtX[i][j] = mean(b_i b_j), XtY[i] = mean(y b_i)
inline void doubleRegressionNormalEquations(const std::vector<RV>& basis, const RV& y, std::vector<double>& XtX, std::vector<double>& XtY) {
    const int k = (int)basis.size();
    std::vector<std::shared_ptr<const RandomVariableFromDoubleArray>> b;
    for (int i = 0; i < k; i++) b.push_back(RandomVariableFromDoubleArray::as_self(basis[(size_t)i]));
    auto cy = RandomVariableFromDoubleArray::as_self(y);
    const int64_t n = cy->size();
    auto at = [&](int i, int64_t p) { return b[(size_t)i]->isDeterministic() ? b[(size_t)i]->doubleValue() : b[(size_t)i]->vec()[(size_t)p]; };
    XtX.assign((size_t)k * k, 0.0); XtY.assign((size_t)k, 0.0);
    for (int i = 0; i < k; i++) {
        for (int j = i; j < k; j++) {
            double s = 0.0;
            for (int64_t p = 0; p < n; p++) s += at(i, p) * at(j, p);
            XtX[(size_t)i * k + j] = XtX[(size_t)j * k + i] = s / (double)n;
        }
        double s = 0.0;
        for (int64_t p = 0; p < n; p++) s += at(i, p) * cy->vec()[(size_t)p];
        XtY[(size_t)i] = s / (double)n;
    }
}

}  // namespace finmath
