// brownian_kernel.cu — Brownian increments on device: MT19937 with polynomial jump-ahead + AS241 inverse normal.
//
// Reproduces, bit for bit, the uniform stream that finmath-lib's BrownianMotionFromMersenneRandomNumbers feeds to
// the reference (call sites: LIBORMarketModelCalibrationATMTest.java:283, MonteCarloBlackScholesModelTest.java:78-85):
// commons-math3 MersenneTwister (== mt19937ar), nextDouble() = (next(26) << 26 | next(26)) * 2^-52, consumed
// path-major (path, then time index, then factor), mapped through Wichura's AS241 PPND16 and scaled by sqrt(dt).
// The reference's own GPU generator (cuRAND XORWOW, BrownianMotionCudaWithRandomVariableCuda.java:155-176) is NOT
// reproduced: it yields a different stream from the CPU path, so CPU/GPU runs there never agree path by path.
//
// Parallelisation: the stream is cut into chunks of 2^15 words. A block starts from the generator state at a
// chunk boundary, obtained from the seeded state by applying the precomputed jump polynomials t^(2^m) mod phi
// (mt_jump_gen.c) for the set bits of the chunk index:  x[k+J] = XOR_{j: c_j = 1} x[k+j].
// Inside a block the 624-word state is regenerated in three data-parallel phases (227 + 227 + 170 words).
// Increments are staged in shared memory as [row = t*F+f][path] tiles and written out as contiguous runs.
//
// Bound: fp64 arithmetic of AS241 (no FMA, to match Java) — not HBM. Algorithmic bytes: 4 per increment written.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace fmc {

namespace {

constexpr int MT_M = 397;
constexpr int BTHREADS = 320;                 // 10 warps: 312 uniforms per regeneration
constexpr int JTHREADS = 640;                 // jump kernel: 624 state words
constexpr int JUMP_DEG = 19937;
constexpr int JUMP_SEQ = JUMP_DEG + MT_N;     // words of the sequence needed to jump one state

__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// One regeneration: nw[0..623] = the 624 words following old[0..623]. All threads of the block must call it.
__device__ __forceinline__ void regenerate(const uint32_t* __restrict__ old, uint32_t* __restrict__ nw, int tid) {
    if (tid < MT_N - MT_M) nw[tid] = twist(old[tid], old[tid + 1], old[tid + MT_M]);
    __syncthreads();
    if (tid < MT_N - MT_M) { const int i = tid + (MT_N - MT_M); nw[i] = twist(old[i], old[i + 1], nw[i - (MT_N - MT_M)]); }
    __syncthreads();
    { const int i = tid + 2 * (MT_N - MT_M);
      if (i < MT_N) nw[i] = twist(old[i], (i + 1 < MT_N) ? old[i + 1] : nw[0], nw[i - (MT_N - MT_M)]); }
    __syncthreads();
}

// Wichura AS241 PPND16 in double, without fused multiply-add (Java semantics); same constants as oracle/fm_oracle.c
__device__ __forceinline__ double mad(double a, double r, double c) { return __dadd_rn(__dmul_rn(a, r), c); }

// The coefficients live in constant memory: a double literal in the code costs two UMOV instructions per use (the kernel is bound
// by instruction issue), an operand in a constant bank costs none.
__constant__ double c_a[8] = {3.3871328727963666080e+00, 1.3314166789178437745e+02, 1.9715909503065514427e+03, 1.3731693765509461125e+04,
                              4.5921953931549871457e+04, 6.7265770927008700853e+04, 3.3430575583588128105e+04, 2.5090809287301226727e+03};
__constant__ double c_b[8] = {1.0, 4.2313330701600911252e+01, 6.8718700749205790830e+02, 5.3941960214247511077e+03, 2.1213794301586595867e+04,
                              3.9307895800092710610e+04, 2.8729085735721942674e+04, 5.2264952788528545610e+03};

// central region |p - 0.5| <= 0.425 (85 % of the uniforms): one rational function, one division
__device__ __forceinline__ double icdf_central(double q) {
    const double r = __dsub_rn(0.180625, __dmul_rn(q, q));
    double num = c_a[7], den = c_b[7];
#pragma unroll
    for (int i = 6; i >= 0; i--) { num = mad(num, r, c_a[i]); den = mad(den, r, c_b[i]); }
    return __ddiv_rn(__dmul_rn(q, num), den);
}

// the tails: log, sqrt and another rational function; kept out of line (the generator runs it on a compacted queue)
__device__ __noinline__ double icdf_tail(double p) {
    const double q = p - 0.5;
    double r = (q < 0.0) ? p : 1.0 - p;
    if (r <= 0.0) return (q < 0.0) ? -INFINITY : INFINITY;
    r = sqrt(-log(r));
    double val;
    if (r <= 5.0) {
        r -= 1.6;
        double num = 7.74545014278341407640e-04;
        num = mad(num, r, 2.27238449892691845833e-02); num = mad(num, r, 2.41780725177450611770e-01);
        num = mad(num, r, 1.27045825245236838258e+00); num = mad(num, r, 3.64784832476320460504e+00);
        num = mad(num, r, 5.76949722146069140550e+00); num = mad(num, r, 4.63033784615654529590e+00);
        num = mad(num, r, 1.42343711074968357734e+00);
        double den = 1.05075007164441684324e-09;
        den = mad(den, r, 5.47593808499534494600e-04); den = mad(den, r, 1.51986665636164571966e-02);
        den = mad(den, r, 1.48103976427480074590e-01); den = mad(den, r, 6.89767334985100004550e-01);
        den = mad(den, r, 1.67638483018380384940e+00); den = mad(den, r, 2.05319162663775882187e+00);
        den = mad(den, r, 1.0);
        val = __ddiv_rn(num, den);
    } else {
        r -= 5.0;
        double num = 2.01033439929228813265e-07;
        num = mad(num, r, 2.71155556874348757815e-05); num = mad(num, r, 1.24266094738807843860e-03);
        num = mad(num, r, 2.65321895265761230930e-02); num = mad(num, r, 2.96560571828504891230e-01);
        num = mad(num, r, 1.78482653991729133580e+00); num = mad(num, r, 5.46378491116411436990e+00);
        num = mad(num, r, 6.65790464350110377720e+00);
        double den = 2.04426310338993978564e-15;
        den = mad(den, r, 1.42151175831644588870e-07); den = mad(den, r, 1.84631831751005468180e-05);
        den = mad(den, r, 7.86869131145613259100e-04); den = mad(den, r, 1.48753612908506148525e-02);
        den = mad(den, r, 1.36929880922735805310e-01); den = mad(den, r, 5.99832206555887937690e-01);
        den = mad(den, r, 1.0);
        val = __ddiv_rn(num, den);
    }
    return (q < 0.0) ? -val : val;
}

// ---------------------------------------------------------------------------------------------------------
// jump kernel: out_states[b] = state at chunk index chunk_of_block[b], starting from the seeded state
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JTHREADS)
mt_jump_kernel(const uint32_t* __restrict__ base_state, const uint32_t* __restrict__ polys /* [npoly][624] */,
               int npoly, const long long* __restrict__ chunk_of_block, uint32_t* __restrict__ out_states)
{
    extern __shared__ uint32_t seq[];            // [JUMP_SEQ + pad] word sequence generated from the current state
    __shared__ uint32_t poly[MT_N];
    const int tid = threadIdx.x;
    long long c = chunk_of_block[blockIdx.x];
    for (int i = tid; i < MT_N; i += JTHREADS) seq[i] = base_state[i];
    __syncthreads();
    for (int m = 0; m < npoly && c != 0; m++, c >>= 1) {
        if (!(c & 1)) continue;
        // extend: seq[624 .. JUMP_SEQ) from seq[0..623]
        for (int base = 0; base + MT_N < JUMP_SEQ + MT_N; base += MT_N) regenerate(seq + base, seq + base + MT_N, tid);
        for (int i = tid; i < MT_N; i += JTHREADS) poly[i] = polys[(long long)m * MT_N + i];
        __syncthreads();
        uint32_t acc = 0;
        if (tid < MT_N) {
            for (int w = 0; w < MT_N; w++) {
                uint32_t bits = poly[w];
                const uint32_t* s = seq + tid + w * 32;
                while (bits) {
                    const int j = __ffs(bits) - 1;
                    bits &= bits - 1;
                    acc ^= s[j];
                }
            }
        }
        __syncthreads();
        if (tid < MT_N) seq[tid] = acc;
        __syncthreads();
    }
    for (int i = tid; i < MT_N; i += JTHREADS) out_states[(long long)blockIdx.x * MT_N + i] = seq[i];
}

// ---------------------------------------------------------------------------------------------------------
// raw words
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BTHREADS)
mt_raw_kernel(const uint32_t* __restrict__ block_states, const long long* __restrict__ chunk_of_block,
              long long words_per_block, unsigned long long skip, long long count, uint32_t* __restrict__ out)
{
    __shared__ uint32_t st[2][MT_N];
    const int tid = threadIdx.x;
    for (int i = tid; i < MT_N; i += BTHREADS) st[0][i] = block_states[(long long)blockIdx.x * MT_N + i];
    __syncthreads();
    // this block emits stream words [w_lo, w_hi)
    const unsigned long long w_lo = skip + (unsigned long long)blockIdx.x * (unsigned long long)words_per_block;
    unsigned long long w_hi = w_lo + (unsigned long long)words_per_block;
    if (w_hi > skip + (unsigned long long)count) w_hi = skip + (unsigned long long)count;
    unsigned long long w = (unsigned long long)chunk_of_block[blockIdx.x] * MT_CHUNK_WORDS;   // stream index of the next regenerated word
    int cur = 0;
    while (w < w_hi) {
        regenerate(st[cur], st[cur ^ 1], tid);
        cur ^= 1;
        for (int i = tid; i < MT_N; i += BTHREADS) {
            const unsigned long long k = w + (unsigned long long)i;
            if (k >= w_lo && k < w_hi) out[k - skip] = temper(st[cur][i]);
        }
        w += MT_N;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Brownian increments
// ---------------------------------------------------------------------------------------------------------
struct BrownianLaunch {
    const uint32_t* block_states;      // [grid][624]
    const long long* chunk_of_block;   // [grid]
    long long paths_per_block;
    long long p0, np;                  // slice [p0, p0+np)
    int T, F;
    int PT;                            // paths per smem tile (odd)
    const double* sqrt_dt;
    float* const* out;                 // [T*F] -> np floats
};

// Tail uniforms (|u - 0.5| > 0.425: 15 % of them, but 99 % of all warps hold at least one) are not evaluated where they occur —
// the whole warp would walk the log / sqrt / second-rational branch for a few lanes — but queued in shared memory with their
// destination and evaluated densely, a block's worth at a time (and always before the tile they belong to is written out).
constexpr int QCAP = 2 * BTHREADS;            // a full round waiting + one more regeneration's tails

__global__ void __launch_bounds__(BTHREADS)
brownian_kernel(const BrownianLaunch P)
{
    extern __shared__ float tiles[];             // [2][TF * PT]
    __shared__ uint32_t st[2][MT_N];
    __shared__ double q_u[QCAP];
    __shared__ uint32_t q_dst[QCAP];             // tile slot (low 16 bits: float index in `tiles` / 1 ... see below) and time index
    __shared__ int q_n;
    const int tid = threadIdx.x;
    const int TF = P.T * P.F;
    const int PT = P.PT;
    const long long tile_elems = (long long)TF * PT;

    const long long pb0 = P.p0 + (long long)blockIdx.x * P.paths_per_block;     // first path of this block
    long long pb1 = pb0 + P.paths_per_block;
    if (pb1 > P.p0 + P.np) pb1 = P.p0 + P.np;
    if (pb0 >= pb1) return;
    const long long e_lo = pb0 * TF, e_hi = pb1 * TF;                            // global element range (1 element = 2 words)

    for (int i = tid; i < MT_N; i += BTHREADS) st[0][i] = P.block_states[(long long)blockIdx.x * MT_N + i];
    if (tid == 0) q_n = 0;
    __syncthreads();
    long long e = P.chunk_of_block[blockIdx.x] * (long long)(MT_CHUNK_WORDS / 2); // element index of the next regeneration's first pair
    int cur = 0;
    long long next_tile = 0;                     // next tile (index within block) to flush
    const long long n_tiles = (pb1 - pb0 + PT - 1) / PT;

    // evaluates queued tails: all of them (force) or full rounds of BTHREADS only. Called by all threads.
    auto drain = [&](bool force) {
        for (;;) {
            __syncthreads();
            const int n = q_n;
            if (n == 0 || (!force && n < BTHREADS)) break;
            const int take = n < BTHREADS ? n : BTHREADS;
            if (tid < take) {
                const int k = n - take + tid;
                const uint32_t d = q_dst[k];
                const double z = icdf_tail(q_u[k]);
                tiles[d & 0xffffu] = __double2float_rn(__dmul_rn(z, P.sqrt_dt[d >> 16]));
            }
            __syncthreads();
            if (tid == 0) q_n = n - take;
        }
    };

    // Where this thread's element lands is tracked incrementally (the kernel was bound by instruction issue: three runtime
    // integer divisions and 64-bit index arithmetic per element cost more than the inverse normal): an element advances by
    // MT_N / 2 per regeneration, i.e. by dq paths and dr rows.
    const int D = MT_N / 2;
    const int dq = D / TF, dr = D - dq * TF;
    const uint32_t magicF = (uint32_t)((0x100000000ull + (unsigned)P.F - 1ull) / (unsigned)P.F);   // row / F == umulhi(row, magicF) for row < 2^16
    bool tracking = false;
    int row = 0, plt = 0;                        // row = t * F + f, path within its tile
    unsigned tpar = 0;                           // parity of the tile index
    long long next_flush_e = 0;                  // a tile is complete once e has passed this element index
    {
        long long tp1 = pb0 + PT; if (tp1 > pb1) tp1 = pb1;
        next_flush_e = tp1 * TF;
    }
    while (e < e_hi) {
        regenerate(st[cur], st[cur ^ 1], tid);
        cur ^= 1;
        const long long my = e + tid;
        if (tid < D && my >= e_lo) {
            if (!tracking) {                     // once per thread: the only divisions
                const unsigned rel = (unsigned)(my - e_lo);       // block-relative element (e_lo is a multiple of TF)
                const unsigned pl = rel / (unsigned)TF;
                row = (int)(rel - pl * (unsigned)TF);
                const unsigned ti = pl / (unsigned)PT;
                plt = (int)(pl - ti * (unsigned)PT);
                tpar = ti & 1u;
                tracking = true;
            }
            if (my < e_hi) {
                const uint32_t hi = temper(st[cur][2 * tid]) >> 6, lo = temper(st[cur][2 * tid + 1]) >> 6;
                const double u = (double)(((unsigned long long)hi << 26) | (unsigned long long)lo) * 0x1.0p-52;
                const int t = (int)__umulhi((unsigned)row, magicF);
                const uint32_t slot = (uint32_t)((int)tpar * (int)tile_elems + row * PT + plt);
                const double q = u - 0.5;
                if (fabs(q) <= 0.425) {
                    tiles[slot] = __double2float_rn(__dmul_rn(icdf_central(q), P.sqrt_dt[t]));
                } else {
                    const int k = atomicAdd(&q_n, 1);
                    q_u[k] = u;
                    q_dst[k] = slot | ((uint32_t)t << 16);
                }
            }
            // advance to this thread's element of the next regeneration
            row += dr; plt += dq;
            if (row >= TF) { row -= TF; plt++; }
            if (plt >= PT) { plt -= PT; tpar ^= 1u; }
            if (plt >= PT) { plt -= PT; tpar ^= 1u; }
        }
        e += D;
        // a tile that is now complete is written out — after the tails still queued for it
        drain(next_tile < n_tiles && next_flush_e <= e);
        while (next_tile < n_tiles && next_flush_e <= e) {
            __syncthreads();
            const long long tp0 = pb0 + next_tile * PT;
            long long tp1 = tp0 + PT; if (tp1 > pb1) tp1 = pb1;
            const int npaths = (int)(tp1 - tp0);
            const float* src = tiles + (next_tile & 1) * tile_elems;
            const int warp = tid >> 5, lane = tid & 31;
            for (int r = warp; r < TF; r += BTHREADS / 32) {
                float* dst = P.out[r] + (tp0 - P.p0);
                const float* sr = src + r * PT;
                for (int j = lane; j < npaths; j += 32) dst[j] = sr[j];
            }
            __syncthreads();
            next_tile++;
            long long tq1 = pb0 + (next_tile + 1) * PT; if (tq1 > pb1) tq1 = pb1;
            next_flush_e = tq1 * TF;
        }
    }
}

}  // namespace

cudaError_t launch_mt_jump(const uint32_t* base_state, const uint32_t* polys, int npoly, const long long* chunk_of_block,
                           uint32_t* out_states, int n_blocks, cudaStream_t stream) {
    const size_t smem = sizeof(uint32_t) * (size_t)(JUMP_SEQ + 2 * MT_N);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    mt_jump_kernel<<<n_blocks, JTHREADS, smem, stream>>>(base_state, polys, npoly, chunk_of_block, out_states);
    return cudaGetLastError();
}

cudaError_t launch_mt_raw(const uint32_t* block_states, const long long* chunk_of_block, int n_blocks, long long words_per_block,
                          unsigned long long skip, long long count, uint32_t* out, cudaStream_t stream) {
    mt_raw_kernel<<<n_blocks, BTHREADS, 0, stream>>>(block_states, chunk_of_block, words_per_block, skip, count, out);
    return cudaGetLastError();
}

int brownian_max_blocks_per_sm(int T, int F, int PT) {
    const size_t smem = sizeof(float) * 2 * (size_t)T * F * PT;
    if (cudaFuncSetAttribute(brownian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, brownian_kernel, BTHREADS, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

cudaError_t launch_brownian(const BrownianParams& B, cudaStream_t stream) {
    BrownianLaunch P;
    P.block_states = B.block_states; P.chunk_of_block = B.chunk_of_block; P.paths_per_block = B.paths_per_block;
    P.p0 = B.p0; P.np = B.np; P.T = B.T; P.F = B.F; P.PT = B.PT; P.sqrt_dt = B.sqrt_dt; P.out = B.out;
    const size_t smem = sizeof(float) * 2 * (size_t)B.T * B.F * B.PT;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(brownian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    brownian_kernel<<<B.n_blocks, BTHREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace fmc
