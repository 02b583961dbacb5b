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

#include "kernels.h"

namespace fmc {

namespace {

constexpr int RTHREADS = 256;
constexpr int RU_MAX = 2;                        // float4 loads per vector per thread and iteration (1 for k > 4: registers)
constexpr int RTILE = RTHREADS * RU_MAX * 4;     // 2048 paths per block iteration

struct Sum {                                     // double sum of exact products + float sum of the float-rounding errors
    double d; float r;
    __device__ __forceinline__ void add(float a, float b, double A, double B) {
        const float p = __fmul_rn(a, b);
        r = __fadd_rn(r, __fmaf_rn(a, b, -p));
        d = __fma_rn(A, B, d);
    }
    __device__ __forceinline__ double value() const { return d - (double)r; }
};

template <int K>
__global__ void __launch_bounds__(RTHREADS)
regression_kernel(const __grid_constant__ RegressionParams P)
{
    constexpr int M = K * (K + 1) / 2 + K;
    constexpr int RU = K <= 4 ? RU_MAX : 1;
    constexpr int SUB = RU_MAX / RU;                 // sub-tiles of RTHREADS * RU * 4 paths per tile
    Sum acc[M];
#pragma unroll
    for (int t = 0; t < M; t++) { acc[t].d = 0.0; acc[t].r = 0.0f; }

    const int tid = threadIdx.x;
    const long long n = P.n;
    const long long n_tiles = (n + RTILE - 1) / RTILE;

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

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if ((tile + 1) * RTILE <= n) {
#pragma unroll
            for (int sub = 0; sub < SUB; sub++) {
                const long long base0 = tile * RTILE + (long long)sub * (RTHREADS * RU * 4) + (long long)tid * 4;
                // full tile: all (K + 1) * RU 128-bit loads of the thread are issued before the first use
                float4 vb[K][RU], vy[RU];
#pragma unroll
                for (int u = 0; u < RU; u++) {
                    const long long base = base0 + (long long)u * RTHREADS * 4;
#pragma unroll
                    for (int i = 0; i < K; i++) {
                        const float* p = P.basis[i];
                        vb[i][u] = p ? __ldg(reinterpret_cast<const float4*>(p + base)) : make_float4(P.scalars[i], P.scalars[i], P.scalars[i], P.scalars[i]);
                    }
                    vy[u] = __ldg(reinterpret_cast<const float4*>(P.y + base));
                }
#pragma unroll
                for (int u = 0; u < RU; u++) {
                    float b[K];
#pragma unroll
                    for (int i = 0; i < K; i++) b[i] = vb[i][u].x;
                    fold(b, vy[u].x);
#pragma unroll
                    for (int i = 0; i < K; i++) b[i] = vb[i][u].y;
                    fold(b, vy[u].y);
#pragma unroll
                    for (int i = 0; i < K; i++) b[i] = vb[i][u].z;
                    fold(b, vy[u].z);
#pragma unroll
                    for (int i = 0; i < K; i++) b[i] = vb[i][u].w;
                    fold(b, vy[u].w);
                }
            }
        } else {
            // ragged last tile: element by element
            for (long long idx = tile * RTILE + tid; idx < n; idx += RTHREADS) {
                float b[K];
#pragma unroll
                for (int i = 0; i < K; i++) b[i] = P.basis[i] ? P.basis[i][idx] : P.scalars[i];
                fold(b, P.y[idx]);
            }
        }
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

template <int K>
cudaError_t launch_k(const RegressionParams& P, int grid, cudaStream_t s) {
    regression_kernel<K><<<grid, RTHREADS, 0, s>>>(P);
    return cudaGetLastError();
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

int regression_max_blocks_per_sm(int k) {
    int nb = 0;
    cudaError_t e;
    switch (k) {
    case 1: case 2: case 3: case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, regression_kernel<4>, RTHREADS, 0); break;
    case 5: case 6: case 7: case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, regression_kernel<8>, RTHREADS, 0); break;
    default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, regression_kernel<12>, RTHREADS, 0); break;
    }
    if (e != cudaSuccess) nb = 1;
    return nb > 0 ? nb : 1;
}
int regression_tile_elems() { return RTILE; }

}  // namespace fmc
