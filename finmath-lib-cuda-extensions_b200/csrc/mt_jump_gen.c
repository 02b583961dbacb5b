/*
 * mt_jump_gen.c — build-time generator of the MT19937 jump-ahead table used by brownian_kernel.cu.
 *
 * The word sequence x[k] of MT19937 satisfies a GF(2)-linear recurrence whose characteristic polynomial phi(t)
 * has degree 19937. For any J:  t^J = sum_j c_j t^j (mod phi)  implies  x[k+J] = XOR_{j: c_j=1} x[k+j]  for all k,
 * so a generator state can be advanced by J words with one pass over 19937+624 consecutive words.
 * This program
 *   1. finds phi with Berlekamp-Massey on 2*19937 output bits,
 *   2. computes P_m = t^(2^m) mod phi for m = LOG2_CHUNK .. LOG2_CHUNK+NPOLY-1 by repeated squaring,
 *   3. self-checks both against the plain recurrence,
 *   4. writes the table as a C initialiser (mt_jump_table.inc): NPOLY x 624 uint32 words, bit j of the table row
 *      = coefficient c_j.
 * The device applies P_m for the set bits of a chunk index to reach any multiple of 2^LOG2_CHUNK words.
 *
 * usage: mt_jump_gen <out.inc> [log2_chunk=15] [npoly=28]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N 624
#define M 397
#define DEG 19937
#define PW 312                 /* 64-bit words holding a polynomial of degree < 19968 */
#define PW2 (2 * PW)

static uint32_t st[N];
static int sti;

static void seed(uint32_t s) {
    st[0] = s;
    for (int i = 1; i < N; i++) st[i] = 1812433253u * (st[i - 1] ^ (st[i - 1] >> 30)) + (uint32_t)i;
    sti = N;
}
static uint32_t next_untempered(void) {
    if (sti >= N) {
        for (int k = 0; k < N; k++) {
            uint32_t y = (st[k] & 0x80000000u) | (st[(k + 1) % N] & 0x7fffffffu);
            st[k] = st[(k + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        sti = 0;
    }
    return st[sti++];
}

typedef uint64_t poly_t[PW2];

static int get_bit(const uint64_t* p, int i) { return (int)((p[i >> 6] >> (i & 63)) & 1u); }
static void flip_bit(uint64_t* p, int i) { p[i >> 6] ^= (uint64_t)1 << (i & 63); }

/* Berlekamp-Massey over GF(2); s has nbits bits; returns L and connection polynomial C (C[0]=1) */
static int berlekamp_massey(const uint64_t* s, int nbits, uint64_t* C /* PW2 words */) {
    static uint64_t B[PW2], Tmp[PW2], srev[2 * PW2 + 2];
    memset(C, 0, sizeof(uint64_t) * PW2);
    memset(B, 0, sizeof(B));
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    /* discrepancy d = sum_{i=0..L} C[i] * s[n-i]; keep a reversed copy of s so the sum is a word-parallel AND */
    memset(srev, 0, sizeof(srev));
    const int total = nbits;
    for (int n = 0; n < nbits; n++) {
        /* reversed index: bit (total-1-n) holds s[n]; window s[n-i] for i=0..L is srev bits [total-1-n, total-1-n+L] */
        if (get_bit(s, n)) flip_bit(srev, total - 1 - n);
        const int base = total - 1 - n;
        uint64_t acc = 0;
        const int words = (L >> 6) + 1;
        const int sh = base & 63, w0 = base >> 6;
        for (int w = 0; w < words; w++) {
            uint64_t v = srev[w0 + w] >> sh;
            if (sh) v |= srev[w0 + w + 1] << (64 - sh);
            acc ^= v & C[w];
        }
        const int d = __builtin_parityll(acc);
        if (d) {
            memcpy(Tmp, C, sizeof(uint64_t) * PW2);
            /* C ^= B << m */
            const int ws = m >> 6, bs = m & 63;
            for (int w = PW2 - 1; w >= ws; w--) {
                uint64_t v = B[w - ws] << bs;
                if (bs && w - ws - 1 >= 0) v |= B[w - ws - 1] >> (64 - bs);
                C[w] ^= v;
            }
            if (2 * L <= n) { L = n + 1 - L; memcpy(B, Tmp, sizeof(uint64_t) * PW2); m = 1; }
            else m++;
        } else m++;
    }
    return L;
}

static uint64_t PHI[PW2];      /* phi(t), bit j = coefficient of t^j, degree DEG */

/* r = a*a mod phi (a of degree < DEG) */
static void sqr_mod(const uint64_t* a, uint64_t* r) {
    static uint64_t wide[PW2 + 1];
    memset(wide, 0, sizeof(wide));
    for (int i = 0; i < DEG; i++) if (get_bit(a, i)) flip_bit(wide, 2 * i);
    for (int i = 2 * DEG - 2; i >= DEG; i--) {
        if (!get_bit(wide, i)) continue;
        const int shift = i - DEG, ws = shift >> 6, bs = shift & 63;
        for (int w = 0; w <= PW; w++) {
            uint64_t v = PHI[w] << bs;
            if (bs && w > 0) v |= PHI[w - 1] >> (64 - bs);
            wide[w + ws] ^= v;
        }
    }
    memcpy(r, wide, sizeof(uint64_t) * PW);
    for (int i = PW; i < PW2; i++) r[i] = 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s out.inc [log2_chunk] [npoly]\n", argv[0]); return 2; }
    const int log2_chunk = argc > 2 ? atoi(argv[2]) : 15;
    const int npoly = argc > 3 ? atoi(argv[3]) : 28;

    /* 1. characteristic polynomial from the least significant bit of 2*DEG consecutive state words */
    static uint64_t bits[PW2 + 2];
    memset(bits, 0, sizeof(bits));
    seed(5489u);
    const int nb = 2 * DEG;
    for (int i = 0; i < nb; i++) if (next_untempered() & 1u) flip_bit(bits, i);
    static uint64_t C[PW2];
    const int L = berlekamp_massey(bits, nb, C);
    if (L != DEG) { fprintf(stderr, "Berlekamp-Massey: linear complexity %d != %d\n", L, DEG); return 1; }
    /* s[n] = sum_{i=1..L} C[i] s[n-i]  <=>  sum_{i=0..L} C[i] E^{L-i} s = 0  =>  phi(t) = sum_i C[i] t^{L-i} */
    memset(PHI, 0, sizeof(PHI));
    for (int i = 0; i <= L; i++) if (get_bit(C, i)) flip_bit(PHI, L - i);

    /* reference word sequence for the self checks */
    const int SEQ = DEG + N + 64;
    uint32_t* x = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(SEQ + (1 << 16) + 8));
    seed(4357u);
    /* x[k] for k >= 0: seeded array first (its word 0 carries 31 unused bits), then generated words */
    for (int i = 0; i < N; i++) x[i] = st[i];
    for (int i = N; i < SEQ + (1 << 16) + 8; i++) x[i] = next_untempered();
    /* check phi annihilates the sequence (k >= 1 so that every term is a clean state word) */
    for (int k = 1; k < 40; k++) {
        uint32_t acc = 0;
        for (int j = 0; j <= DEG; j++) if (get_bit(PHI, j)) acc ^= x[k + j];
        if (acc != 0) { fprintf(stderr, "phi does not annihilate the sequence at k=%d\n", k); return 1; }
    }

    /* 2. P_m = t^(2^m) mod phi */
    static uint64_t P[PW2], Q[PW2];
    memset(P, 0, sizeof(P));
    flip_bit(P, 1);   /* t */
    FILE* f = fopen(argv[1], "w");
    if (!f) { perror(argv[1]); return 1; }
    fprintf(f, "// generated by mt_jump_gen.c: row m = coefficients of t^(2^(%d+m)) mod phi_MT19937, bit j of the row = c_j\n", log2_chunk);
    fprintf(f, "#define MT_JUMP_LOG2_CHUNK %d\n#define MT_JUMP_NPOLY %d\n", log2_chunk, npoly);
    fprintf(f, "static const uint32_t kMtJumpTable[MT_JUMP_NPOLY][624] = {\n");
    for (int m = 1; m < log2_chunk + npoly; m++) {
        sqr_mod(P, Q);
        memcpy(P, Q, sizeof(P));
        if (m == 16) {
            /* 3. self check: x[k + 2^16] == XOR_j c_j x[k+j] */
            for (int k = 0; k < N; k++) {
                uint32_t acc = 0;
                for (int j = 0; j < DEG; j++) if (get_bit(P, j)) acc ^= x[k + j];
                const uint32_t want = x[k + (1 << 16)];
                const uint32_t mask = (k == 0) ? 0x80000000u : 0xffffffffu;   /* only the top bit of the oldest word is state */
                if ((acc ^ want) & mask) { fprintf(stderr, "jump self-check failed at k=%d\n", k); return 1; }
            }
        }
        if (m >= log2_chunk) {
            fprintf(f, "  {");
            for (int w = 0; w < N; w++) {
                const uint32_t v = (uint32_t)(P[w >> 1] >> ((w & 1) * 32));
                fprintf(f, "0x%08xu%s", v, w + 1 < N ? "," : "");
                if ((w & 7) == 7) fprintf(f, "\n   ");
            }
            fprintf(f, "}%s\n", m + 1 < log2_chunk + npoly ? "," : "");
        }
    }
    fprintf(f, "};\n");
    fclose(f);
    free(x);
    fprintf(stderr, "mt_jump_gen: wrote %s (%d polynomials from 2^%d)\n", argv[1], npoly, log2_chunk);
    return 0;
}
