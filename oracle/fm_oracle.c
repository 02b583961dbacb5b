/*
 * fm_oracle.c — CPU ORACLE (test infrastructure, NOT product code). See fm_oracle.h for scope and pinning status.
 *
 * Every function cites the reference line it restates:
 *   RVF = /root/reference/src/main/java/net/finmath/cuda/cpu/montecarlo/RandomVariableFromFloatArray.java
 * Code that restates un-vendored dependencies (commons-math3 MersenneTwister, finmath-lib NormalDistribution /
 * BrownianMotionFromMersenneRandomNumbers / MonteCarloConditionalExpectationRegression) follows the published
 * algorithms (mt19937ar.c, Wichura AS241) and the reference's call sites
 * (LIBORMarketModelCalibrationATMTest.java:283, MonteCarloBlackScholesModelTest.java:78-85,
 *  BrownianMotionJavaRandom.java:156-178 for the "uniform -> ICDF * sqrt(dt)" recipe).
 *
 * Compile with -ffp-contract=off: Java never contracts a*b+c into an FMA.
 */
#include "fm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------- Java semantics helpers ---------- */

/* java.lang.Math.min(float,float) / FastMath.min: NaN if either is NaN; -0.0f < +0.0f (RVF:759, 1167) */
static inline float java_minf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return (signbit(a) || signbit(b)) ? -0.0f : 0.0f;
    return a < b ? a : b;
}
static inline float java_maxf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return (signbit(a) && signbit(b)) ? -0.0f : 0.0f;
    return a > b ? a : b;
}
static inline double java_min(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (signbit(a) || signbit(b)) ? -0.0 : 0.0;
    return a < b ? a : b;
}
static inline double java_max(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (signbit(a) && signbit(b)) ? -0.0 : 0.0;
    return a > b ? a : b;
}
/* java.lang.Math.pow differs from C99 pow in two corners: pow(x, NaN) is NaN even for x == 1, and
 * pow(+-1, +-inf) is NaN. */
static inline double java_pow(double x, double y) {
    if (y != y) return y;
    if (y == 0.0) return 1.0;
    if (x != x) return x;
    if (isinf(y) && fabs(x) == 1.0) return NAN;
    return pow(x, y);
}
/* java.lang.Math.round(double) -> long: floor(x + 0.5) (RVF:484) */
static inline long java_round(double x) { return (long)floor(x + 0.5); }

/* ---------- conversions ---------- */

void orc_from_f64(const double* in, float* out, int64_t n) {            /* RVF:217-223 */
    for (int64_t i = 0; i < n; i++) out[i] = (float)in[i];
}
void orc_to_f64(const float* in, double* out, int64_t n) {              /* RVF:225-231 */
    for (int64_t i = 0; i < n; i++) out[i] = in[i];
}

/* ---------- elementwise ---------- */

int orc_op_vs(int op, const float* x, double s, float* out, int64_t n) {
    const float fs = (float)s;
    switch (op) {
    case ORC_CAP:   for (int64_t i = 0; i < n; i++) out[i] = java_minf(x[i], fs); return 0;          /* RVF:757-760 */
    case ORC_FLOOR: for (int64_t i = 0; i < n; i++) out[i] = java_maxf(x[i], fs); return 0;          /* RVF:772-775 */
    case ORC_ADD:   for (int64_t i = 0; i < n; i++) out[i] = x[i] + fs; return 0;                    /* RVF:787-790 */
    case ORC_SUB:   for (int64_t i = 0; i < n; i++) out[i] = x[i] - fs; return 0;                    /* RVF:802-805 */
    case ORC_BUS:   for (int64_t i = 0; i < n; i++) out[i] = fs - x[i]; return 0;                    /* RVF:1033-1038 with roles swapped */
    case ORC_MULT:  for (int64_t i = 0; i < n; i++) out[i] = x[i] * fs; return 0;                    /* RVF:817-820 */
    case ORC_DIV:   for (int64_t i = 0; i < n; i++) out[i] = x[i] / fs; return 0;                    /* RVF:832-835 */
    case ORC_VID:   for (int64_t i = 0; i < n; i++) out[i] = fs / x[i]; return 0;                    /* RVF:1098-1103 with roles swapped */
    case ORC_POW:   for (int64_t i = 0; i < n; i++) out[i] = (float)java_pow((double)x[i], (double)fs); return 0; /* RVF:847-850 */
    default: return -1;
    }
}

int orc_op_v(int op, const float* x, float* out, int64_t n) {
    switch (op) {
    case ORC_SQUARED: for (int64_t i = 0; i < n; i++) out[i] = x[i] * x[i]; return 0;                /* RVF:873-876 */
    case ORC_SQRT:    for (int64_t i = 0; i < n; i++) out[i] = (float)sqrt((double)x[i]); return 0;  /* RVF:888-891 */
    case ORC_EXP:     for (int64_t i = 0; i < n; i++) out[i] = (float)exp((double)x[i]); return 0;   /* RVF:903-906 */
    case ORC_LOG:     for (int64_t i = 0; i < n; i++) out[i] = (float)log((double)x[i]); return 0;   /* RVF:918-921 */
    case ORC_SIN:     for (int64_t i = 0; i < n; i++) out[i] = (float)sin((double)x[i]); return 0;   /* RVF:933-936 */
    case ORC_COS:     for (int64_t i = 0; i < n; i++) out[i] = (float)cos((double)x[i]); return 0;   /* RVF:948-951 */
    case ORC_INVERT:  for (int64_t i = 0; i < n; i++) out[i] = 1.0f / x[i]; return 0;                /* RVF:1294-1297 */
    case ORC_ABS:     for (int64_t i = 0; i < n; i++) out[i] = fabsf(x[i]); return 0;                /* RVF:1309-1312 */
    case ORC_ISNAN:   for (int64_t i = 0; i < n; i++) out[i] = (x[i] != x[i]) ? 1.0f : 0.0f; return 0; /* RVF:1445-1448 */
    default: return -1;
    }
}

int orc_op_vv(int op, const float* x, const float* y, float* out, int64_t n) {
    switch (op) {
    case ORC_ADD:   for (int64_t i = 0; i < n; i++) out[i] = x[i] + y[i]; return 0;                  /* RVF:981-984 */
    case ORC_SUB:   for (int64_t i = 0; i < n; i++) out[i] = x[i] - y[i]; return 0;                  /* RVF:1011-1014 */
    case ORC_BUS:   for (int64_t i = 0; i < n; i++) out[i] = y[i] - x[i]; return 0;                  /* RVF:1041-1044 */
    case ORC_MULT:  for (int64_t i = 0; i < n; i++) out[i] = x[i] * y[i]; return 0;                  /* RVF:1073-1076 */
    case ORC_DIV:   for (int64_t i = 0; i < n; i++) out[i] = x[i] / y[i]; return 0;                  /* RVF:1106-1109 */
    case ORC_VID:   for (int64_t i = 0; i < n; i++) out[i] = (float)((double)y[i] / (double)x[i]); return 0; /* RVF:1136-1139: double division, then round */
    case ORC_CAP:   for (int64_t i = 0; i < n; i++) out[i] = java_minf(x[i], y[i]); return 0;        /* RVF:1165-1168 */
    case ORC_FLOOR: for (int64_t i = 0; i < n; i++) out[i] = java_maxf(x[i], y[i]); return 0;        /* RVF:1194-1197 */
    default: return -1;
    }
}

int orc_op_vvs(int op, const float* x, const float* y, double s, float* out, int64_t n) {
    const float fs = (float)s;
    switch (op) {
    case ORC_ACCRUE:     for (int64_t i = 0; i < n; i++) out[i] = x[i] * (1.0f + y[i] * fs); return 0; /* RVF:1222-1225 */
    case ORC_DISCOUNT:   for (int64_t i = 0; i < n; i++) out[i] = x[i] / (1.0f + y[i] * fs); return 0; /* RVF:1250-1253 */
    case ORC_ADDPRODUCT: for (int64_t i = 0; i < n; i++) out[i] = x[i] + y[i] * fs; return 0;          /* RVF:1345-1348 */
    default: return -1;
    }
}

int orc_op_vvv(int op, const float* x, const float* y, const float* z, float* out, int64_t n) {
    switch (op) {
    case ORC_ADDPRODUCT: for (int64_t i = 0; i < n; i++) out[i] = x[i] + y[i] * z[i]; return 0;       /* RVF:1374-1377 */
    case ORC_CHOOSE:     for (int64_t i = 0; i < n; i++) out[i] = (float)((double)x[i] >= 0.0 ? (double)y[i] : (double)z[i]); return 0; /* RVF:1280-1282 */
    case ORC_ADDRATIO:   for (int64_t i = 0; i < n; i++) out[i] = x[i] + y[i] / z[i]; return 0;       /* RVF:1409-1412 */
    case ORC_SUBRATIO:   for (int64_t i = 0; i < n; i++) out[i] = x[i] - y[i] / z[i]; return 0;       /* RVF:1432-1435 */
    default: return -1;
    }
}

/* ---------- reductions ---------- */

double orc_min(const float* x, int64_t n) {                               /* RVF:284-296 */
    double m = 1.7976931348623157e308;
    if (n != 0) m = x[0];
    for (int64_t i = 0; i < n; i++) m = java_min((double)x[i], m);
    return m;
}
double orc_max(const float* x, int64_t n) {                               /* RVF:299-311 */
    double m = -1.7976931348623157e308;
    if (n != 0) m = x[0];
    for (int64_t i = 0; i < n; i++) m = java_max((double)x[i], m);
    return m;
}
double orc_average(const float* x, int64_t n) {                           /* RVF:314-334 */
    if (n == 0) return NAN;
    double sum = 0.0, error = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double value = (double)x[i] - error;
        const double newSum = sum + value;
        error = (newSum - sum) - value;
        sum = newSum;
    }
    return sum / (double)n;
}
double orc_average_f64(const double* x, int64_t n) {
    if (n == 0) return NAN;
    double sum = 0.0, error = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double value = x[i] - error;
        const double newSum = sum + value;
        error = (newSum - sum) - value;
        sum = newSum;
    }
    return sum / (double)n;
}
double orc_average_w(const float* x, const float* prob, int64_t n) {     /* RVF:337-357 */
    if (n == 0) return NAN;
    double sum = 0.0, error = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double value = (double)x[i] * (double)prob[i] - error;
        const double newSum = sum + value;
        error = (newSum - sum) - value;
        sum = newSum;
    }
    return sum / (double)n;
}
double orc_variance(const float* x, int64_t n) {                          /* RVF:360-382 */
    if (n == 1) return 0.0;
    if (n == 0) return NAN;
    const double average = orc_average(x, n);
    double sum = 0.0, errorOfSum = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double value = ((double)x[i] - average) * ((double)x[i] - average) - errorOfSum;
        const double newSum = sum + value;
        errorOfSum = (newSum - sum) - value;
        sum = newSum;
    }
    return sum / (double)n;
}
double orc_variance_w(const float* x, const float* prob, int64_t n) {    /* RVF:385-407 */
    if (n == 0) return NAN;
    const double average = orc_average_w(x, prob, n);
    double sum = 0.0, errorOfSum = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const double value = ((double)x[i] - average) * ((double)x[i] - average) * (double)prob[i] - errorOfSum;
        const double newSum = sum + value;
        errorOfSum = (newSum - sum) - value;
        sum = newSum;
    }
    return sum;
}
double orc_sample_variance(const float* x, int64_t n) {                   /* RVF:410-419 */
    if (n == 1) return 0.0;
    if (n == 0) return NAN;
    return orc_variance(x, n) * (double)n / (double)(n - 1);
}

static int cmp_float(const void* a, const void* b) {
    /* java.util.Arrays.sort(float[]): total order, -0.0f < 0.0f, NaN last */
    const float fa = *(const float*)a, fb = *(const float*)b;
    if (fa != fa) return (fb != fb) ? 0 : 1;
    if (fb != fb) return -1;
    if (fa < fb) return -1;
    if (fa > fb) return 1;
    if (fa == 0.0f && fb == 0.0f) return (int)signbit(fb) - (int)signbit(fa);
    return 0;
}
static float* sorted_copy(const float* x, int64_t n) {
    float* s = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    memcpy(s, x, sizeof(float) * (size_t)n);
    qsort(s, (size_t)n, sizeof(float), cmp_float);
    return s;
}
static int64_t quantile_index(int64_t n, double q) {                      /* RVF:484 */
    long idx = java_round((double)(n + 1) * q - 1.0);
    if (idx < 0) idx = 0;
    if (idx > n - 1) idx = n - 1;
    return idx;
}
double orc_quantile(const float* x, int64_t n, double q) {                /* RVF:473-487 */
    if (n == 0) return NAN;
    float* s = sorted_copy(x, n);
    const double r = s[quantile_index(n, q)];
    free(s);
    return r;
}
double orc_quantile_expectation(const float* x, int64_t n, double q0, double q1) { /* RVF:502-526 */
    if (n == 0) return NAN;
    if (q0 > q1) return orc_quantile_expectation(x, n, q1, q0);
    float* s = sorted_copy(x, n);
    const int64_t i0 = quantile_index(n, q0), i1 = quantile_index(n, q1);
    double e = 0.0;
    for (int64_t i = i0; i <= i1; i++) e += s[i];
    e /= (double)(i1 - i0 + 1);
    free(s);
    return e;
}
void orc_histogram(const float* x, int64_t n, const double* pts, int m, double* out) { /* RVF:548-578 */
    float* s = sorted_copy(x, n);
    int64_t sampleIndex = 0;
    for (int k = 0; k < m; k++) {
        int64_t count = 0;
        while (sampleIndex < n && (double)s[sampleIndex] <= pts[k]) { sampleIndex++; count++; }
        out[k] = (double)count;
    }
    out[m] = (double)(n - sampleIndex);
    if (n > 0) for (int k = 0; k <= m; k++) out[k] /= (double)n;
    free(s);
}

/* ---------- MT19937 (commons-math3 MersenneTwister == mt19937ar.c) ---------- */

#define MT_N 624
#define MT_M 397

void orc_mt_seed_int(orc_mt_t* g, uint32_t seed) {                        /* MersenneTwister.setSeed(int) */
    g->mt[0] = seed;
    for (int i = 1; i < MT_N; i++)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->mti = MT_N;
}
void orc_mt_seed_array(orc_mt_t* g, const uint32_t* key, int len) {       /* MersenneTwister.setSeed(int[]) */
    orc_mt_seed_int(g, 19650218u);
    int i = 1, j = 0;
    for (int k = (MT_N > len ? MT_N : len); k != 0; k--) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= MT_N) { g->mt[0] = g->mt[MT_N - 1]; i = 1; }
        if (j >= len) j = 0;
    }
    for (int k = MT_N - 1; k != 0; k--) {
        g->mt[i] = (g->mt[i] ^ ((g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= MT_N) { g->mt[0] = g->mt[MT_N - 1]; i = 1; }
    }
    g->mt[0] = 0x80000000u;
    g->mti = MT_N;
}
void orc_mt_seed_long(orc_mt_t* g, int64_t seed) {                        /* MersenneTwister.setSeed(long) */
    uint32_t key[2] = { (uint32_t)((uint64_t)seed >> 32), (uint32_t)((uint64_t)seed & 0xffffffffu) };
    orc_mt_seed_array(g, key, 2);
}
uint32_t orc_mt_next_u32(orc_mt_t* g) {                                   /* MersenneTwister.next(32) */
    static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
    uint32_t y;
    if (g->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
            g->mt[kk] = g->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
            g->mt[kk] = g->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (g->mt[MT_N - 1] & 0x80000000u) | (g->mt[0] & 0x7fffffffu);
        g->mt[MT_N - 1] = g->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        g->mti = 0;
    }
    y = g->mt[g->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
double orc_mt_next_double(orc_mt_t* g) {                                  /* BitsStreamGenerator.nextDouble */
    const uint64_t high = ((uint64_t)(orc_mt_next_u32(g) >> 6)) << 26;
    const uint64_t low = (uint64_t)(orc_mt_next_u32(g) >> 6);
    return (double)(high | low) * 0x1.0p-52;
}
static void seed_by_mode(orc_mt_t* g, int seed_mode, int64_t seed) {
    if (seed_mode == 1) orc_mt_seed_int(g, (uint32_t)seed);
    else orc_mt_seed_long(g, seed);
}
void orc_mt_fill_u32(int seed_mode, int64_t seed, uint64_t skip, uint32_t* out, int64_t count) {
    orc_mt_t g;
    seed_by_mode(&g, seed_mode, seed);
    for (uint64_t i = 0; i < skip; i++) (void)orc_mt_next_u32(&g);
    for (int64_t i = 0; i < count; i++) out[i] = orc_mt_next_u32(&g);
}

/* ---------- Wichura AS241 PPND16 (finmath-lib NormalDistribution.inverseCumulativeDistribution) ---------- */

double orc_icdf(double p) {
    static const double a0 = 3.3871328727963666080e+00, a1 = 1.3314166789178437745e+02, a2 = 1.9715909503065514427e+03,
                        a3 = 1.3731693765509461125e+04, a4 = 4.5921953931549871457e+04, a5 = 6.7265770927008700853e+04,
                        a6 = 3.3430575583588128105e+04, a7 = 2.5090809287301226727e+03;
    static const double b1 = 4.2313330701600911252e+01, b2 = 6.8718700749205790830e+02, b3 = 5.3941960214247511077e+03,
                        b4 = 2.1213794301586595867e+04, b5 = 3.9307895800092710610e+04, b6 = 2.8729085735721942674e+04,
                        b7 = 5.2264952788528545610e+03;
    static const double c0 = 1.42343711074968357734e+00, c1 = 4.63033784615654529590e+00, c2 = 5.76949722146069140550e+00,
                        c3 = 3.64784832476320460504e+00, c4 = 1.27045825245236838258e+00, c5 = 2.41780725177450611770e-01,
                        c6 = 2.27238449892691845833e-02, c7 = 7.74545014278341407640e-04;
    static const double d1 = 2.05319162663775882187e+00, d2 = 1.67638483018380384940e+00, d3 = 6.89767334985100004550e-01,
                        d4 = 1.48103976427480074590e-01, d5 = 1.51986665636164571966e-02, d6 = 5.47593808499534494600e-04,
                        d7 = 1.05075007164441684324e-09;
    static const double e0 = 6.65790464350110377720e+00, e1 = 5.46378491116411436990e+00, e2 = 1.78482653991729133580e+00,
                        e3 = 2.96560571828504891230e-01, e4 = 2.65321895265761230930e-02, e5 = 1.24266094738807843860e-03,
                        e6 = 2.71155556874348757815e-05, e7 = 2.01033439929228813265e-07;
    static const double f1 = 5.99832206555887937690e-01, f2 = 1.36929880922735805310e-01, f3 = 1.48753612908506148525e-02,
                        f4 = 7.86869131145613259100e-04, f5 = 1.84631831751005468180e-05, f6 = 1.42151175831644588870e-07,
                        f7 = 2.04426310338993978564e-15;
    const double split1 = 0.425, split2 = 5.0, const1 = 0.180625, const2 = 1.6;

    const double q = p - 0.5;
    double r, val;
    if (fabs(q) <= split1) {
        r = const1 - q * q;
        return q * (((((((a7 * r + a6) * r + a5) * r + a4) * r + a3) * r + a2) * r + a1) * r + a0) /
                   (((((((b7 * r + b6) * r + b5) * r + b4) * r + b3) * r + b2) * r + b1) * r + 1.0);
    }
    r = (q < 0.0) ? p : 1.0 - p;
    if (r <= 0.0) return (q < 0.0) ? -INFINITY : INFINITY;   /* p <= 0 / p >= 1; unreachable for u in [2^-52, 1-2^-52] */
    r = sqrt(-log(r));
    if (r <= split2) {
        r -= const2;
        val = (((((((c7 * r + c6) * r + c5) * r + c4) * r + c3) * r + c2) * r + c1) * r + c0) /
              (((((((d7 * r + d6) * r + d5) * r + d4) * r + d3) * r + d2) * r + d1) * r + 1.0);
    } else {
        r -= split2;
        val = (((((((e7 * r + e6) * r + e5) * r + e4) * r + e3) * r + e2) * r + e1) * r + e0) /
              (((((((f7 * r + f6) * r + f5) * r + f4) * r + f3) * r + f2) * r + f1) * r + 1.0);
    }
    return (q < 0.0) ? -val : val;
}
void orc_icdf_array(const double* p, double* out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = orc_icdf(p[i]);
}

/* ---------- Brownian increments ---------- */

void orc_brownian_f64(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1, const double* sqrt_dt, double* out) {
    orc_mt_t g;
    seed_by_mode(&g, seed_mode, seed);
    const int64_t np = p1 - p0;
    const uint64_t skip = 2ull * (uint64_t)T * (uint64_t)F * (uint64_t)p0;
    for (uint64_t i = 0; i < skip; i++) (void)orc_mt_next_u32(&g);
    for (int64_t p = 0; p < np; p++)
        for (int t = 0; t < T; t++)
            for (int f = 0; f < F; f++) {
                const double u = orc_mt_next_double(&g);
                out[((int64_t)t * F + f) * np + p] = orc_icdf(u) * sqrt_dt[t];
            }
}
void orc_brownian(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1, const double* sqrt_dt, float* out) {
    orc_mt_t g;
    seed_by_mode(&g, seed_mode, seed);
    const int64_t np = p1 - p0;
    const uint64_t skip = 2ull * (uint64_t)T * (uint64_t)F * (uint64_t)p0;
    for (uint64_t i = 0; i < skip; i++) (void)orc_mt_next_u32(&g);
    for (int64_t p = 0; p < np; p++)
        for (int t = 0; t < T; t++)
            for (int f = 0; f < F; f++) {
                const double u = orc_mt_next_double(&g);
                out[((int64_t)t * F + f) * np + p] = (float)(orc_icdf(u) * sqrt_dt[t]);   /* RVF:217-223 cast at wrap time */
            }
}

/* ---------- regression normal equations ---------- */

static double avg_prod(const float* a, double sa, const float* b, double sb, int64_t n, float* tmp) {
    /* RVF.mult dispatch (RVF:1050-1079) followed by getAverage (RVF:314-334) */
    if (!a && !b) return sa * sb;                                            /* RVF:1059-1061 */
    if (a && b) { for (int64_t i = 0; i < n; i++) tmp[i] = a[i] * b[i]; }    /* RVF:1073-1076 */
    else if (a) { const float f = (float)sb; for (int64_t i = 0; i < n; i++) tmp[i] = a[i] * f; } /* RVF:1063-1064 -> 817-820 */
    else        { const float f = (float)sa; for (int64_t i = 0; i < n; i++) tmp[i] = f * b[i]; } /* RVF:1066-1069 */
    return orc_average(tmp, n);
}
void orc_regression_normal_eq(const float* const* basis, const double* scalars, int k, const float* y, int64_t n,
                              double* XtX, double* Xty) {
    float* tmp = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < k; i++) {
        for (int j = i; j < k; j++) {
            const double v = avg_prod(basis[i], scalars[i], basis[j], scalars[j], n, tmp);
            XtX[i * k + j] = v;
            XtX[j * k + i] = v;
        }
        Xty[i] = avg_prod(y, 0.0, basis[i], scalars[i], n, tmp);
    }
    free(tmp);
}
