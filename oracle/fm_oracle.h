/*
 * fm_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's CPU float vector path
 *   src/main/java/net/finmath/cuda/cpu/montecarlo/RandomVariableFromFloatArray.java  ("RVF")
 * plus the pieces of finmath-lib 5.1.3 / commons-math3 3.6.1 that the hot path depends on and that are
 * NOT vendored under /root/reference (MersenneTwister, BitsStreamGenerator.nextDouble,
 * NormalDistribution.inverseCumulativeDistribution = Wichura AS241 PPND16,
 * BrownianMotionFromMersenneRandomNumbers loop order, MonteCarloConditionalExpectationRegression normal equations).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this library.
 * The product (libfmcuda.so) never links, loads or calls it.
 *
 * Pinning status:
 *   - elementwise + reductions: pinned by the known-answer vectors of RandomVariableGPUTest.java:68-188 (tests/test_oracle_kat.py)
 *   - MT19937 uint32 stream: pinned by the mt19937ar known answers (init_genrand(5489), init_by_array{0x123,0x234,0x345,0x456})
 *     and cross-checked word by word against numpy.random.MT19937.
 *   - AS241 inverse normal: cross-checked against scipy.special.ndtri (different algorithm, agreement to ~1e-15).
 *   - seeding mode used by finmath-lib's BrownianMotionFromMersenneRandomNumbers (int vs long seed), regression
 *     coefficients, LMM prices: PARITY UNPINNED by the reference (no golden vectors in /root/reference, no JVM here).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math  (Java float semantics: IEEE binary32/64, RN, no FMA contraction).
 */
#ifndef FM_ORACLE_H
#define FM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- opcodes (shared numbering with include/fmcuda.h; the numbers are part of the test contract only) ---- */
enum {
    /* vector (op) scalar — RVF:751-853 */
    ORC_CAP = 1, ORC_FLOOR = 2, ORC_ADD = 3, ORC_SUB = 4, ORC_BUS = 5, ORC_MULT = 6, ORC_DIV = 7, ORC_VID = 8, ORC_POW = 9,
    /* unary — RVF:867-954, 1287-1315, 1440-1451 */
    ORC_SQUARED = 20, ORC_SQRT = 21, ORC_EXP = 22, ORC_LOG = 23, ORC_SIN = 24, ORC_COS = 25, ORC_INVERT = 26, ORC_ABS = 27, ORC_ISNAN = 28,
    /* ternary — RVF:1202-1256, 1263-1285, 1317-1438 */
    ORC_ACCRUE = 40, ORC_DISCOUNT = 41, ORC_ADDPRODUCT = 42, ORC_CHOOSE = 43, ORC_ADDRATIO = 44, ORC_SUBRATIO = 45
};

/* RVF:217-223 getFloatArray: (float)d[i] */
void orc_from_f64(const double* in, float* out, int64_t n);
/* RVF:225-231 getDoubleArray */
void orc_to_f64(const float* in, double* out, int64_t n);

/* x (op) (float)s    — stochastic branch of RVF:751-853 (cap floor add sub mult div pow) and the scalar-this
 * branches of RVF:974-1197 for bus/vid: bus = (float)s - x, vid = (float)s / x. returns 0 or -1 on bad opcode */
int orc_op_vs(int op, const float* x, double s, float* out, int64_t n);
/* unary */
int orc_op_v(int op, const float* x, float* out, int64_t n);
/* x (op) y, both stochastic — last branch of RVF:960-1200 (add sub bus mult div vid cap floor) */
int orc_op_vv(int op, const float* x, const float* y, float* out, int64_t n);
/* accrue / discount (x, rate, p) RVF:1221-1226,1249-1254; addProduct(x, f1, (float)s) RVF:1344-1349 */
int orc_op_vvs(int op, const float* x, const float* y, double s, float* out, int64_t n);
/* addProduct(x,f1,f2) RVF:1373-1378; choose(trigger,a,b) RVF:1277-1284; addRatio/subRatio RVF:1408-1413,1431-1436 */
int orc_op_vvv(int op, const float* x, const float* y, const float* z, float* out, int64_t n);

/* reductions, RVF:283-470 */
double orc_min(const float* x, int64_t n);
double orc_max(const float* x, int64_t n);
double orc_average(const float* x, int64_t n);                         /* RVF:314-334 Kahan in double */
double orc_average_w(const float* x, const float* prob, int64_t n);    /* RVF:337-357 */
double orc_variance(const float* x, int64_t n);                        /* RVF:360-382 */
double orc_variance_w(const float* x, const float* prob, int64_t n);   /* RVF:385-407 (NOT divided by n) */
double orc_sample_variance(const float* x, int64_t n);                 /* RVF:410-419 */
double orc_quantile(const float* x, int64_t n, double q);              /* RVF:473-487 */
double orc_quantile_expectation(const float* x, int64_t n, double q0, double q1); /* RVF:502-526 */
void   orc_histogram(const float* x, int64_t n, const double* interval_points, int m, double* out /* m+1 */); /* RVF:529-581 */

/* double-storage twin (finmath-lib RandomVariableFromDoubleArray; same formulas in double). Used for the
 * "RandomVariableFromDoubleArray path" CPU baseline only. */
double orc_average_f64(const double* x, int64_t n);

/* ---- MT19937 as in commons-math3 3.6.1 org.apache.commons.math3.random.MersenneTwister ---- */
typedef struct { uint32_t mt[624]; int mti; } orc_mt_t;
void     orc_mt_seed_int(orc_mt_t* g, uint32_t seed);                  /* setSeed(int)  == init_genrand */
void     orc_mt_seed_array(orc_mt_t* g, const uint32_t* key, int len); /* setSeed(int[]) == init_by_array */
void     orc_mt_seed_long(orc_mt_t* g, int64_t seed);                  /* setSeed(long) == init_by_array{hi32, lo32} */
uint32_t orc_mt_next_u32(orc_mt_t* g);                                 /* next(32): tempered output */
double   orc_mt_next_double(orc_mt_t* g);                              /* BitsStreamGenerator.nextDouble: (next(26)<<26 | next(26)) * 2^-52 */
void     orc_mt_fill_u32(int seed_mode, int64_t seed, uint64_t skip, uint32_t* out, int64_t count);

/* finmath-lib NormalDistribution.inverseCumulativeDistribution: Wichura AS241 PPND16 in double, no FMA */
double orc_icdf(double p);
void   orc_icdf_array(const double* p, double* out, int64_t n);

/* finmath-lib BrownianMotionFromMersenneRandomNumbers.doGenerateBrownianMotion (recalled):
 * loop path (outer) -> timeIndex -> factor; inc = icdf(mt.nextDouble()) * sqrt_dt[t]; stored (float) per (t,f) vector.
 * seed_mode 0 = setSeed(long) (default of finmath 5.x wrapper), 1 = setSeed(int).
 * Generates paths [p0, p1) of a motion with n_total paths into out[(t*F+f)*(p1-p0) + (p-p0)]. */
void orc_brownian(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1, const double* sqrt_dt, float* out);
/* same in double storage (RandomVariableFromDoubleArray twin) */
void orc_brownian_f64(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1, const double* sqrt_dt, double* out);

/* finmath-lib MonteCarloConditionalExpectationRegression normal equations (recalled):
 * XtX[i][j] = average( (float)(b_i * b_j) ), Xty[i] = average( (float)(y * b_i) ); basis[i]==NULL means the
 * deterministic constant scalars[i] (then b_i*b_j follows the deterministic/stochastic dispatch of RVF.mult). */
void orc_regression_normal_eq(const float* const* basis, const double* scalars, int k, const float* y, int64_t n,
                              double* XtX /* k*k */, double* Xty /* k */);

#ifdef __cplusplus
}
#endif
#endif
