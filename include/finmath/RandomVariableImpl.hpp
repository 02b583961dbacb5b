// RandomVariableImpl.hpp — the host-side object model of a RandomVariable backed by a vector backend.
//
// This is the C++ statement of the dispatch logic of RandomVariableCuda (RVC,
// /root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:1171-1704) with the semantics of
// its CPU twin RandomVariableFromFloatArray (RVF, .../cuda/cpu/montecarlo/RandomVariableFromFloatArray.java:751-1451)
// where RVC is incomplete or defective:
//   - stochastic = backend vector + filtration time; deterministic = one double, size() == 1 (RVC:566-577);
//   - deterministic (op) deterministic in double on the host; a deterministic operand of a stochastic vector is passed to
//     the backend as a scalar (cast to float there, RVC:521);
//   - an operand with a higher type priority takes over the operation (mirror methods per RVF:962-1178);
//   - filtration time of a result = max of the operand times (RVF:968).
// The backend policy B supplies the vector arithmetic: CudaBackend (RandomVariableCuda.hpp, the product, C ABI of
// include/fmcuda.h) or the CPU oracle's FloatArrayBackend (oracle/RandomVariableFromFloatArray.hpp, tests/baseline only).
#pragma once
#include <algorithm>
#include <cmath>
#include <memory>
#include <typeinfo>
#include <utility>

#include "RandomVariable.hpp"

namespace finmath {

// opcodes shared by all backends (numbering of include/fmcuda.h and oracle/fm_oracle.h)
enum : int {
    OP_CAP = 1, OP_FLOOR = 2, OP_ADD = 3, OP_SUB = 4, OP_BUS = 5, OP_MULT = 6, OP_DIV = 7, OP_VID = 8, OP_POW = 9,
    OP_SQUARED = 20, OP_SQRT = 21, OP_EXP = 22, OP_LOG = 23, OP_SIN = 24, OP_COS = 25, OP_INVERT = 26, OP_ABS = 27, OP_ISNAN = 28,
    OP_ACCRUE = 40, OP_DISCOUNT = 41, OP_ADDPRODUCT = 42, OP_CHOOSE = 43, OP_ADDRATIO = 44, OP_SUBRATIO = 45
};
enum : int { R_SUM = 1, R_AVERAGE = 2, R_VARIANCE = 3, R_SAMPLE_VARIANCE = 4, R_MIN = 5, R_MAX = 6, R_AVERAGE_W = 7, R_VARIANCE_W = 8 };

namespace detail {
inline double java_min(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (std::signbit(a) || std::signbit(b)) ? -0.0 : 0.0;
    return a < b ? a : b;
}
inline double java_max(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (std::signbit(a) && std::signbit(b)) ? -0.0 : 0.0;
    return a > b ? a : b;
}
inline double java_pow(double x, double y) {
    if (y != y) return y;
    if (y == 0.0) return 1.0;
    if (x != x) return x;
    if (std::isinf(y) && std::fabs(x) == 1.0) return NAN;
    return std::pow(x, y);
}
// Every operation returns a new RandomVariable object: one heap block (object + shared_ptr control block) per recorded operation,
// 50 000 per LMM step. The blocks are recycled through a per-thread free list instead of malloc / free (a block freed by another
// thread joins that thread's list; the list is bounded).
template <class T>
struct RecycleAlloc {
    using value_type = T;
    RecycleAlloc() = default;
    template <class U> RecycleAlloc(const RecycleAlloc<U>&) noexcept {}
    struct Cache {
        void* head = nullptr; size_t count = 0;
        ~Cache() { while (head) { void* next = *static_cast<void**>(head); ::operator delete(head); head = next; } count = 0; }
    };
    static Cache& cache() { static thread_local Cache c; return c; }
    T* allocate(std::size_t n) {
        static_assert(sizeof(T) >= sizeof(void*), "a free block holds the link to the next one");
        if (n == 1) {
            Cache& c = cache();
            if (c.head) { void* p = c.head; c.head = *static_cast<void**>(p); c.count--; return static_cast<T*>(p); }
        }
        return static_cast<T*>(::operator new(n * sizeof(T)));
    }
    void deallocate(T* p, std::size_t n) noexcept {
        if (n == 1) {
            Cache& c = cache();
            if (c.count < (1u << 16)) { *reinterpret_cast<void**>(p) = c.head; c.head = p; c.count++; return; }
        }
        ::operator delete(p);
    }
    template <class U> bool operator==(const RecycleAlloc<U>&) const noexcept { return true; }
    template <class U> bool operator!=(const RecycleAlloc<U>&) const noexcept { return false; }
};
}  // namespace detail

template <class B>
class RandomVariableImpl final : public RandomVariable {
public:
    using Vec = typename B::Vec;
    using Self = RandomVariableImpl<B>;

    // deterministic
    RandomVariableImpl(double time, double value, int typePriority = B::kTypePriority)
        : time_(time), det_(true), value_(value), n_(1), priority_(typePriority) {}
    // stochastic, takes ownership of the backend vector
    RandomVariableImpl(double time, Vec&& vec, int64_t n, int typePriority = B::kTypePriority)
        : time_(time), det_(false), value_(NAN), vec_(std::move(vec)), n_(n), priority_(typePriority) {}
    ~RandomVariableImpl() override { if (!det_) B::release(vec_); }
    RandomVariableImpl(const RandomVariableImpl&) = delete;
    RandomVariableImpl& operator=(const RandomVariableImpl&) = delete;

    static RV of(double time, double value) { return std::allocate_shared<Self>(detail::RecycleAlloc<Self>(), time, value); }                       // RVC:643-646
    static RV of(double time, Vec&& vec, int64_t n) { return std::allocate_shared<Self>(detail::RecycleAlloc<Self>(), time, std::move(vec), n); }    // RVC:631-634
    static RV of(double time, const double* values, int64_t n) { return of(time, B::from_f64(values, n), n); }     // RVC:723-725

    const Vec& vec() const { return vec_; }

    // ---- accessors ----
    double getFiltrationTime() const override { return time_; }
    int getTypePriority() const override { return priority_; }
    bool isDeterministic() const override { return det_; }
    int64_t size() const override { return det_ ? 1 : n_; }
    double get(int64_t i) const override { return det_ ? value_ : B::get(vec_, n_, i); }
    double doubleValue() const override {
        if (!det_) throw UnsupportedOperationException("The random variable is non-deterministic");              // RVC:1124-1131
        return value_;
    }
    std::vector<double> getRealizations() const override {                                                         // RVC:1114-1122
        if (det_) return {value_};
        return B::to_f64(vec_, n_);
    }

    // ---- statistics (RVF:283-526) ----
    double getMin() const override { return det_ ? value_ : B::reduce(R_MIN, vec_, n_, nullptr); }
    double getMax() const override { return det_ ? value_ : B::reduce(R_MAX, vec_, n_, nullptr); }
    double getAverage() const override { return det_ ? value_ : B::reduce(R_AVERAGE, vec_, n_, nullptr); }
    double getAverage(const RV& p) const override {
        if (det_) return value_ * p->getAverage();                                                                 // RVF:338-340
        auto q = ref(p);
        if (q->det_) return this->mult(q->value_)->getAverage();
        return B::reduce(R_AVERAGE_W, vec_, n_, &q->vec_);
    }
    double getVariance() const override { return det_ ? 0.0 : B::reduce(R_VARIANCE, vec_, n_, nullptr); }
    double getVariance(const RV& p) const override {
        if (det_) return 0.0;
        auto q = as_self(p);
        if (q->det_) { std::vector<double> w((size_t)n_, q->value_); q = cast(of(q->time_, w.data(), n_)); }
        return B::reduce(R_VARIANCE_W, vec_, n_, &q->vec_);
    }
    double getSampleVariance() const override { return det_ ? 0.0 : B::reduce(R_SAMPLE_VARIANCE, vec_, n_, nullptr); }
    double getQuantile(double q) const override { return det_ ? value_ : B::quantile(vec_, n_, q); }
    double getQuantileExpectation(double q0, double q1) const override { return det_ ? value_ : B::quantile_expectation(vec_, n_, q0, q1); }

    // ---- scalar / unary operators (RVC:1171-1352) ----
    RV cap(double c) const override { return det_ ? det(detail::java_min(value_, c)) : vs(OP_CAP, c); }
    RV floor(double f) const override { return det_ ? det(detail::java_max(value_, f)) : vs(OP_FLOOR, f); }
    RV add(double v) const override { return det_ ? det(value_ + v) : vs(OP_ADD, v); }
    RV sub(double v) const override { return det_ ? det(value_ - v) : vs(OP_SUB, v); }
    RV bus(double v) const override { return det_ ? det(-value_ + v) : vs(OP_BUS, v); }
    RV mult(double v) const override { return det_ ? det(value_ * v) : vs(OP_MULT, v); }
    RV div(double v) const override { return det_ ? det(value_ / v) : vs(OP_DIV, v); }
    RV vid(double v) const override { return det_ ? det(v / value_) : vs(OP_VID, v); }
    RV pow(double e) const override { return det_ ? det(detail::java_pow(value_, e)) : vs(OP_POW, e); }
    RV average() const override { return of(-1.7976931348623157e308, getAverage()); }                              // RVC:1279-1282
    RV squared() const override { return det_ ? det(value_ * value_) : v(OP_SQUARED); }
    RV sqrt() const override { return det_ ? det(std::sqrt(value_)) : v(OP_SQRT); }
    RV exp() const override { return det_ ? det(std::exp(value_)) : v(OP_EXP); }
    RV log() const override { return det_ ? det(std::log(value_)) : v(OP_LOG); }
    RV sin() const override { return det_ ? det(std::sin(value_)) : v(OP_SIN); }                                   // RVF:926-939
    RV cos() const override { return det_ ? det(std::cos(value_)) : v(OP_COS); }                                   // RVF:941-954
    RV invert() const override { return det_ ? det(1.0 / value_) : v(OP_INVERT); }
    RV abs() const override { return det_ ? det(std::fabs(value_)) : v(OP_ABS); }
    RV isNaN() const override { return det_ ? det(value_ != value_ ? 1.0 : 0.0) : v(OP_ISNAN); }                   // RVF:1440-1451

    // ---- binary operators with type priority (RVC:1390-1580; mirrors per RVF) ----
    RV add(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->add(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, value_ + r->doubleValue());
        if (det_) return ref(r)->vs(OP_ADD, value_, t);                                                        // RVF:974-979
        if (r->isDeterministic()) return vs(OP_ADD, r->doubleValue(), t);
        return vv(OP_ADD, *ref(r), t);
    }
    RV sub(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->bus(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, value_ - r->doubleValue());
        if (det_) return ref(r)->vs(OP_BUS, value_, t);                                                        // RVF:1003-1008
        if (r->isDeterministic()) return vs(OP_SUB, r->doubleValue(), t);
        return vv(OP_SUB, *ref(r), t);
    }
    RV bus(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->sub(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, -value_ + r->doubleValue());
        if (det_) return ref(r)->vs(OP_SUB, value_, t);                                                        // RVF:1033-1038
        if (r->isDeterministic()) return vs(OP_BUS, r->doubleValue(), t);
        return vv(OP_BUS, *ref(r), t);
    }
    RV mult(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->mult(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, value_ * r->doubleValue());
        if (r->isDeterministic()) return vs(OP_MULT, r->doubleValue(), t);
        if (det_) return ref(r)->vs(OP_MULT, value_, t);                                                       // RVF:1065-1070
        return vv(OP_MULT, *ref(r), t);
    }
    RV div(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->vid(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, value_ / r->doubleValue());
        if (det_) return ref(r)->vs(OP_VID, value_, t);                                                        // RVF:1098-1103
        if (r->isDeterministic()) return vs(OP_DIV, r->doubleValue(), t);
        return vv(OP_DIV, *ref(r), t);
    }
    RV vid(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->div(self());                                               // RVF:1116-1119
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, r->doubleValue() / value_);
        if (det_) return ref(r)->vs(OP_DIV, value_, t);                                                        // RVF:1128-1133
        if (r->isDeterministic()) return vs(OP_VID, r->doubleValue(), t);
        return vv(OP_VID, *ref(r), t);
    }
    RV cap(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->cap(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, detail::java_min(value_, r->doubleValue()));
        if (det_) return ref(r)->vs(OP_CAP, value_, t);                                                        // RVF:1158-1163
        if (r->isDeterministic()) return vs(OP_CAP, r->doubleValue(), t);                                          // missing in RVC:1546-1555 (NPE)
        return vv(OP_CAP, *ref(r), t);
    }
    RV floor(const RV& r) const override {
        if (r->getTypePriority() > priority_) return r->floor(self());
        const double t = std::max(time_, r->getFiltrationTime());
        if (det_ && r->isDeterministic()) return of(t, detail::java_max(value_, r->doubleValue()));
        if (det_) return ref(r)->vs(OP_FLOOR, value_, t);                                                      // RVF:1187-1192
        if (r->isDeterministic()) return vs(OP_FLOOR, r->doubleValue(), t);
        return vv(OP_FLOOR, *ref(r), t);
    }

    // ---- accrue / discount (RVC:1582-1624, RVF:1202-1256) ----
    RV accrue(const RV& rate, double p) const override {
        if (rate->getTypePriority() > priority_) return rate->mult(p)->add(1.0)->mult(self());
        const double t = std::max(time_, rate->getFiltrationTime());
        if (rate->isDeterministic()) return mult(1.0 + rate->doubleValue() * p);
        auto r = ref(rate);
        if (det_) return cast(cast(r->vs(OP_MULT, p))->vs(OP_ADD, 1.0))->vs(OP_MULT, value_, t);                   // RVF:1214-1219
        return of(t, B::vvs(OP_ACCRUE, vec_, r->vec_, p, n_), n_);
    }
    RV discount(const RV& rate, double p) const override {
        if (rate->getTypePriority() > priority_) return rate->mult(p)->add(1.0)->vid(self());                      // RVF:1232-1235
        const double t = std::max(time_, rate->getFiltrationTime());
        if (rate->isDeterministic()) return div(1.0 + rate->doubleValue() * p);
        auto r = ref(rate);
        if (det_) return cast(cast(r->vs(OP_MULT, p))->vs(OP_ADD, 1.0))->vs(OP_VID, value_, t);                    // RVF:1242-1247, RVC:1614-1618
        return of(t, B::vvs(OP_DISCOUNT, vec_, r->vec_, p, n_), n_);
    }

    // ---- ternary (RVF:1263-1438) ----
    RV choose(const RV& a, const RV& b) const override {
        const double t = std::max(std::max(time_, a->getFiltrationTime()), b->getFiltrationTime());
        if (det_) return value_ >= 0 ? a : b;                                                                      // RVF:1270-1276
        auto ca = ref(a), cb = ref(b);
        return of(t, B::choose(vec_, ca->det_ ? nullptr : &ca->vec_, ca->value_, cb->det_ ? nullptr : &cb->vec_, cb->value_, n_), n_);
    }
    RV addProduct(const RV& f1, double f2) const override {
        if (f1->getTypePriority() > priority_) return f1->mult(f2)->add(self());
        const double t = std::max(time_, f1->getFiltrationTime());
        if (f1->isDeterministic()) return add(f1->doubleValue() * f2);
        auto c1 = ref(f1);
        if (det_) return cast(c1->vs(OP_MULT, f2))->vs(OP_ADD, value_, t);                                         // RVF:1329-1334
        return of(t, B::vvs(OP_ADDPRODUCT, vec_, c1->vec_, f2, n_), n_);
    }
    RV addProduct(const RV& f1, const RV& f2) const override {
        if (f1->getTypePriority() > priority_ || f2->getTypePriority() > priority_) return f1->mult(f2)->add(self());
        const double t = std::max(std::max(time_, f1->getFiltrationTime()), f2->getFiltrationTime());
        if (det_ && f1->isDeterministic() && f2->isDeterministic()) return of(t, value_ + f1->doubleValue() * f2->doubleValue());
        if (f1->isDeterministic() && f2->isDeterministic()) return add(f1->doubleValue() * f2->doubleValue());
        if (f2->isDeterministic()) return addProduct(f1, f2->doubleValue());
        if (f1->isDeterministic()) return addProduct(f2, f1->doubleValue());
        if (!det_) {
            auto c1 = ref(f1), c2 = ref(f2);
            return of(t, B::vvv(OP_ADDPRODUCT, vec_, c1->vec_, c2->vec_, n_), n_);
        }
        return add(f1->mult(f2));                                                                                  // RVF:1379-1381
    }
    RV addRatio(const RV& num, const RV& den) const override {
        if (num->getTypePriority() > priority_ || den->getTypePriority() > priority_) return num->div(den)->add(self());   // RVF:1396-1399
        return ratio(num, den, OP_ADDRATIO, +1.0);
    }
    RV subRatio(const RV& num, const RV& den) const override {
        if (num->getTypePriority() > priority_ || den->getTypePriority() > priority_) return num->div(den)->mult(-1.0)->add(self());   // RVF:1419-1422
        return ratio(num, den, OP_SUBRATIO, -1.0);
    }

    // getRandomVariableCuda (RVC:759-766): a foreign RandomVariable is converted (uploaded) on the fly
    static std::shared_ptr<const Self> as_self(const RV& r) {
        if (auto s = std::dynamic_pointer_cast<const Self>(r)) return s;
        if (r->isDeterministic()) return cast(of(r->getFiltrationTime(), r->doubleValue()));
        const std::vector<double> v = r->getRealizations();
        return cast(of(r->getFiltrationTime(), v.data(), (int64_t)v.size()));
    }

    // the same for an operand that is only read during the call: an object of this class is borrowed from the caller's reference
    // (no reference-count traffic, no dynamic_pointer_cast: 100 000 operands per LMM step), a foreign one is converted and kept alive
    struct Ref {
        const Self* p; std::shared_ptr<const Self> keep;
        const Self* operator->() const { return p; }
        const Self& operator*() const { return *p; }
    };
    static Ref ref(const RV& r) {
        const RandomVariable* raw = r.get();
        if (typeid(*raw) == typeid(Self)) return Ref{static_cast<const Self*>(raw), nullptr};
        std::shared_ptr<const Self> s = as_self(r);
        const Self* q = s.get();
        return Ref{q, std::move(s)};
    }

private:
    double time_;
    bool det_;
    double value_;
    Vec vec_{};
    int64_t n_;
    int priority_;

    static std::shared_ptr<const Self> cast(const RV& r) { return std::static_pointer_cast<const Self>(r); }
    RV det(double value) const { return std::allocate_shared<Self>(detail::RecycleAlloc<Self>(), time_, value, priority_); }
    RV vs(int op, double s) const { return of(time_, B::vs(op, vec_, s, n_), n_); }
    RV vs(int op, double s, double t) const { return of(t, B::vs(op, vec_, s, n_), n_); }
    RV v(int op) const { return of(time_, B::v(op, vec_, n_), n_); }
    RV vv(int op, const Self& o, double t) const { return of(t, B::vv(op, vec_, o.vec_, n_, o.n_), n_); }
    RV ratio(const RV& num, const RV& den, int op, double sign) const {
        const double t = std::max(std::max(time_, num->getFiltrationTime()), den->getFiltrationTime());
        if (det_ && num->isDeterministic() && den->isDeterministic()) return of(t, value_ + sign * (num->doubleValue() / den->doubleValue()));
        auto n = ref(num), d = ref(den);
        if (!det_ && !n->det_ && !d->det_) return of(t, B::vvv(op, vec_, n->vec_, d->vec_, n_), n_);
        if (n->det_ && d->det_) {                        // RVF:1408-1413: (float)n / (float)d in float, then added to the vector
            const float q = (float)n->value_ / (float)d->value_;
            return vs(sign > 0 ? OP_ADD : OP_SUB, (double)q, t);
        }
        RV q = n->div(den);
        return sign > 0 ? q->add(self()) : q->bus(self());
    }
};

}  // namespace finmath
