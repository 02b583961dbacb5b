// Exhaustive check (all 2^32 float inputs) of the exp / log sequences of csrc/tape_interp.cuh (f_exp, f_log), restated with the same IEEE
// double operations (and the same table, csrc/log_table.inc), against glibc exp / log rounded to float (the oracle).
// gcc -O2 -mfma -fopenmp -ffp-contract=off explog_exhaustive.c -lm; ./a.out 0 (exp), ./a.out 1 (log): 0 mismatches each.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <omp.h>

static inline double as_double(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t as_u64(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// ---- exp ----
static inline float my_exp(float x) {
    // clamp (NaN handled at the end)
    float xc = fminf(fmaxf(x, -110.0f), 90.0f);
    double xd = (double)xc;
    const double L2E = 1.4426950408889634074, MAGIC = 6755399441055744.0;
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double t = fma(xd, L2E, MAGIC);
    double k = t - MAGIC;
    int32_t ki = (int32_t)(uint32_t)as_u64(t);
    double r = fma(-k, LN2_HI, xd);
    r = fma(-k, LN2_LO, r);
    // Taylor degree 13 (Horner)
    double p = 1.0 / 479001600.0;
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 1.0 / 2.0);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    uint64_t u = as_u64(p) + ((uint64_t)(int64_t)ki << 52);
    float res = (float)as_double(u);
    return x != x ? x + x : res;
}

// ---- log ----
static const double LOGTAB[129][2] = {
#include "../../finmath-lib-cuda-extensions_b200/csrc/log_table.inc"
};
static inline float my_log(float x) {
    double xd = (double)x;
    uint64_t u = as_u64(xd);
    int32_t hi = (int32_t)(u >> 32);
    int32_t e = (hi - 0x3fe6a09e) >> 20;
    int32_t him = hi - (e << 20);
    double m = as_double(((uint64_t)(uint32_t)him << 32) | (u & 0xffffffffu));
    int32_t k = (him >> 13) - (0x3fe6a09e >> 13);
    if (k < 0) k = 0; if (k > 128) k = 128;              // only for garbage inputs (x <= 0, NaN, inf): overridden below
    double invc = LOGTAB[k][0], logc = LOGTAB[k][1];
    double g = fma(m, invc, -1.0);
    double q = -1.0 / 8.0;
    q = fma(q, g, 1.0 / 7.0);
    q = fma(q, g, -1.0 / 6.0);
    q = fma(q, g, 1.0 / 5.0);
    q = fma(q, g, -1.0 / 4.0);
    q = fma(q, g, 1.0 / 3.0);
    q = fma(q, g, -0.5);
    double g2 = g * g;
    double p = fma(g2, q, g);
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    double dk = (double)e;
    double w = fma(dk, ln2_lo, p);
    double w2 = logc + w;
    float r = (float)fma(dk, ln2_hi, w2);
    if (x == 0.0f) r = -INFINITY;
    if (x < 0.0f) r = NAN;
    if (x != x) r = x + x;
    if (x == INFINITY) r = INFINITY;
    return r;
}
int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    long long mism = 0, mism2 = 0, total = 0;
    uint32_t first_bad[8]; int nbad = 0;
#pragma omp parallel for reduction(+:mism, mism2, total) schedule(dynamic, 1)
    for (long long blk = 0; blk < 4096; blk++) {
        for (uint64_t i = (uint64_t)blk << 20; i < ((uint64_t)blk + 1) << 20; i++) {
            float x = as_float((uint32_t)i);
            float want = which == 0 ? (float)exp((double)x) : (float)log((double)x);
            float got = which == 0 ? my_exp(x) : my_log(x);
            total++;
            if (as_u32(want) != as_u32(got) && !(want != want && got != got)) {
                mism++;
                int32_t d = (int32_t)as_u32(want) - (int32_t)as_u32(got);
                if (d > 1 || d < -1) mism2++;
#pragma omp critical
                if (nbad < 8) first_bad[nbad++] = (uint32_t)i;
            }
        }
    }
    printf("%s: total %lld mismatches %lld (more than 1 ulp: %lld)\n", which == 0 ? "exp" : "log", total, mism, mism2);
    for (int i = 0; i < nbad; i++) { float x = as_float(first_bad[i]); printf("  x=%a (%g): want %a got %a\n", x, x, which == 0 ? (float)exp((double)x) : (float)log((double)x), which == 0 ? my_exp(x) : my_log(x)); }
    return 0;
}
