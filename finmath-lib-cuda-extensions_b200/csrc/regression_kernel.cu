// regression_kernel.cu — conditional-expectation regression: the normal equations X^T X and X^T y in ONE pass.
//
// finmath-lib's MonteCarloConditionalExpectationRegression (not vendored; hook in the reference:
// RandomVariableFromFloatArray.java:861-864) builds XtX[i][j] = basis[i].mult(basis[j]).getAverage() and
// XtY[i] = y.mult(basis[i]).getAverage(): k(k+1)/2 + k separate mult + getAverage round trips, each a full
// device->host copy in the reference (RandomVariableCuda.java:869-883). Here every path's k+1 values are read
// once, the products are formed in float (one rounding each, like RVF:1073-1076) and summed with the precision of a
// double sum. k <= 12, so this is a bandwidth-bound reduction (4*(k+1) bytes per path), not a tensor-core contraction.
//
// Accumulation. k = 8 means 44 running sums per path. Widening every float product to double costs one F2F conversion
// per sum per path on a pipe an eighth as wide as the FP32 pipe and made the first version of this kernel
// conversion-bound (28 % of the HBM roofline). Here only the k + 1 INPUTS of a path are widened. With A, B the widened
// operands, A*B is the exact product and one DFMA adds it to the running double sum; the float product the reference
// forms is p = fl32(a*b) = A*B - r with r = fma(a, b, -p) its (exactly representable) rounding error, so the reference's
// sum is  sum(A*B) - sum(r).  The r are ~2^-24 of the terms and are summed in plain float; their own rounding is
// ~n * 2^-48 of the sum of magnitudes for the n terms of one thread. Merging is in double: warp shuffles, warps in order,
// last block over the block partials in a fixed order (deterministic for a given grid).
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>

#include "kernels.h"

namespace fmc {

namespace {

constexpr int RTHREADS = 256;
constexpr int RTILE = RTHREADS * 4;              // 1024 paths per tile: one 128-bit group per thread and vector
constexpr int RTILE_BYTES = RTILE * 4;           // 4 KB per vector and tile
constexpr int RSTAGES_MAX = 6;
constexpr int RSMEM_BUDGET = 200 * 1024;         // dynamic shared memory the ring may take (one CTA per SM)

// Feeding. One thread cannot keep enough loads in flight from registers: 44 running sums (k = 8) leave room for one 128-bit
// load per vector, 28-36 KB per SM in flight against the ~45 KB that HBM latency x bandwidth asks for — the register-fed
// version of this kernel sat at 43-45 % of the roofline for k >= 6 with the SM idle two thirds of the time. Tiles now arrive
// by TMA bulk copies (cp.async.bulk, one per vector and tile, issued by one thread) into a ring of up to 6 stages in shared
// memory (k = 6: 4 stages x 28 KB), each guarded by an mbarrier; the threads read their 128-bit group of every vector from
// the stage, fold it, and a block barrier hands the stage back for the tile RSTAGES ahead.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RDONE;\n"
        "bra RWAIT;\n"
        "RDONE:\n"
        "}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// FLOATP = true: the sum of the FLOAT products fl32(a*b), what RandomVariableFromFloatArray's mult + getAverage give: double sum
// of exact products + float sum of the float-rounding errors (header). FLOATP = false: the sum of the EXACT products — what
// RandomVariableFromDoubleArray (finmath-lib's default CPU type) computes from the same float-valued inputs: one DFMA per
// product and nothing else. The two differ by sum(r), ~2^-24 / sqrt(n) of the sum; the three float operations per product
// that reproduce it make the kernel compute-bound (k = 6: 49 % of the roofline; without them: HBM-bound).
template <bool FLOATP> struct Sum;
template <> struct Sum<true> {
    double d; float r;
    __device__ __forceinline__ void zero() { d = 0.0; r = 0.0f; }
    __device__ __forceinline__ void add(float a, float b, double A, double B) {
        const float p = __fmul_rn(a, b);
        r = __fadd_rn(r, __fmaf_rn(a, b, -p));
        d = __fma_rn(A, B, d);
    }
    __device__ __forceinline__ double value() const { return d - (double)r; }
};
template <> struct Sum<false> {
    double d;
    __device__ __forceinline__ void zero() { d = 0.0; }
    __device__ __forceinline__ void add(float, float, double A, double B) { d = __fma_rn(A, B, d); }
    __device__ __forceinline__ double value() const { return d; }
};

template <int K, bool FLOATP>
__global__ void __launch_bounds__(RTHREADS, 1)
regression_kernel(const __grid_constant__ RegressionParams P, const int n_stages)
{
    constexpr int M = K * (K + 1) / 2 + K;
    extern __shared__ __align__(128) unsigned char ring[];          // [n_stages][K + 1][RTILE_BYTES], then the mbarriers
    Sum<FLOATP> acc[M];
#pragma unroll
    for (int t = 0; t < M; t++) acc[t].zero();

    const int tid = threadIdx.x;
    const long long n = P.n;
    const long long n_full = n / RTILE;                              // full tiles go through the ring
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t stage_bytes = (uint32_t)(K + 1) * RTILE_BYTES;
    const uint32_t mbar0 = ring0 + (uint32_t)n_stages * stage_bytes;
    int n_vec = 1;                                                   // vectors that are really loaded (scalars are not)
#pragma unroll
    for (int i = 0; i < K; i++) n_vec += P.basis[i] ? 1 : 0;

    auto fold = [&](const float (&b)[K], const float y) {
        double B[K];
#pragma unroll
        for (int i = 0; i < K; i++) B[i] = (double)b[i];
        const double Y = (double)y;
        int t = 0;
#pragma unroll
        for (int i = 0; i < K; i++)
#pragma unroll
            for (int j = i; j < K; j++, t++) acc[t].add(b[i], b[j], B[i], B[j]);
#pragma unroll
        for (int i = 0; i < K; i++, t++) acc[t].add(y, b[i], Y, B[i]);
    };
    auto issue = [&](long long tile, int stage) {                    // thread 0: arm the stage's barrier, one bulk copy per vector
        const uint32_t mbar = mbar0 + 8u * (uint32_t)stage;
        const uint32_t dst = ring0 + (uint32_t)stage * stage_bytes;
        mbar_expect_tx(mbar, (uint32_t)n_vec * RTILE_BYTES);
#pragma unroll
        for (int i = 0; i < K; i++)
            if (P.basis[i]) bulk_load(dst + (uint32_t)i * RTILE_BYTES, P.basis[i] + tile * RTILE, RTILE_BYTES, mbar);
        bulk_load(dst + (uint32_t)K * RTILE_BYTES, P.y + tile * RTILE, RTILE_BYTES, mbar);
    };

    if (tid == 0) {
        for (int s = 0; s < n_stages; s++) mbar_init(mbar0 + 8u * (uint32_t)s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < n_stages; s++) {
            const long long tile = blockIdx.x + (long long)s * gridDim.x;
            if (tile < n_full) issue(tile, s);
        }
    }
    __syncthreads();

    int stage = 0;
    uint32_t parity = 0u;
    for (long long tile = blockIdx.x; tile < n_full; tile += gridDim.x) {
        mbar_wait(mbar0 + 8u * (uint32_t)stage, parity);
        const unsigned char* st = ring + (size_t)stage * stage_bytes + (size_t)tid * 16;
        float4 vb[K];
#pragma unroll
        for (int i = 0; i < K; i++)
            vb[i] = P.basis[i] ? *reinterpret_cast<const float4*>(st + (size_t)i * RTILE_BYTES) : make_float4(P.scalars[i], P.scalars[i], P.scalars[i], P.scalars[i]);
        const float4 vy = *reinterpret_cast<const float4*>(st + (size_t)K * RTILE_BYTES);
        __syncthreads();                                             // everybody has read the stage: it may be refilled
        if (tid == 0) {
            const long long next = tile + (long long)n_stages * gridDim.x;
            if (next < n_full) issue(next, stage);
        }
        float b[K];
#pragma unroll
        for (int i = 0; i < K; i++) b[i] = vb[i].x;
        fold(b, vy.x);
#pragma unroll
        for (int i = 0; i < K; i++) b[i] = vb[i].y;
        fold(b, vy.y);
#pragma unroll
        for (int i = 0; i < K; i++) b[i] = vb[i].z;
        fold(b, vy.z);
#pragma unroll
        for (int i = 0; i < K; i++) b[i] = vb[i].w;
        fold(b, vy.w);
        if (++stage == n_stages) { stage = 0; parity ^= 1u; }
    }
    // ragged tail (less than one tile): block 0, element by element
    if (blockIdx.x == 0)
        for (long long idx = n_full * RTILE + tid; idx < n; idx += RTHREADS) {
            float b[K];
#pragma unroll
            for (int i = 0; i < K; i++) b[i] = P.basis[i] ? P.basis[i][idx] : P.scalars[i];
            fold(b, P.y[idx]);
        }

    // block reduction in double: warp shuffle tree, then warps in order (deterministic)
    __shared__ double smem[RTHREADS / 32][M];
    __shared__ bool is_last;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int t = 0; t < M; t++) {
        double v = acc[t].value();
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) smem[warp][t] = v;
    }
    __syncthreads();
    if (tid < M) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < RTHREADS / 32; w++) v += smem[w][tid];
        P.partials[(long long)blockIdx.x * 128 + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned ticket = atomicAdd(P.counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block: entry t is summed over blocks by warp (t % 8), lanes striding the blocks, fixed order
    for (int t = warp; t < M; t += RTHREADS / 32) {
        double v = 0.0;
        for (unsigned k = lane; k < gridDim.x; k += 32) v += ((const volatile double*)P.partials)[(long long)k * 128 + t];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) P.result[t] = v;
    }
    if (tid == 0) *P.counter = 0u;
}

inline int stages_for(int k) { return std::max(2, std::min(RSTAGES_MAX, RSMEM_BUDGET / ((k + 1) * RTILE_BYTES))); }

template <int K, bool FLOATP>
cudaError_t launch_kf(const RegressionParams& P, int grid, cudaStream_t s) {
    const int n_stages = stages_for(K);
    const size_t smem = (size_t)n_stages * (K + 1) * RTILE_BYTES + 8 * RSTAGES_MAX;
    static bool opted = false;                                       // per instantiation
    if (!opted) {
        cudaError_t e = cudaFuncSetAttribute(regression_kernel<K, FLOATP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        opted = true;
    }
    regression_kernel<K, FLOATP><<<grid, RTHREADS, smem, s>>>(P, n_stages);
    return cudaGetLastError();
}
template <int K>
cudaError_t launch_k(const RegressionParams& P, int grid, cudaStream_t s) {
    return P.float_products ? launch_kf<K, true>(P, grid, s) : launch_kf<K, false>(P, grid, s);
}

}  // namespace

cudaError_t launch_regression(const RegressionParams& P, int grid, cudaStream_t stream) {
    switch (P.k) {
    case 1: return launch_k<1>(P, grid, stream);
    case 2: return launch_k<2>(P, grid, stream);
    case 3: return launch_k<3>(P, grid, stream);
    case 4: return launch_k<4>(P, grid, stream);
    case 5: return launch_k<5>(P, grid, stream);
    case 6: return launch_k<6>(P, grid, stream);
    case 7: return launch_k<7>(P, grid, stream);
    case 8: return launch_k<8>(P, grid, stream);
    case 9: return launch_k<9>(P, grid, stream);
    case 10: return launch_k<10>(P, grid, stream);
    case 11: return launch_k<11>(P, grid, stream);
    case 12: return launch_k<12>(P, grid, stream);
    default: return cudaErrorInvalidValue;
    }
}

int regression_max_blocks_per_sm(int) { return 1; }                  // the ring takes most of the SM's shared memory
int regression_tile_elems() { return RTILE; }

}  // namespace fmc
