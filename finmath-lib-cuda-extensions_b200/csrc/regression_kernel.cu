// regression_kernel.cu — conditional-expectation regression: the normal equations X^T X and X^T y in ONE pass.
//
// finmath-lib's MonteCarloConditionalExpectationRegression (not vendored; hook in the reference:
// RandomVariableFromFloatArray.java:861-864) builds XtX[i][j] = basis[i].mult(basis[j]).getAverage() and
// XtY[i] = y.mult(basis[i]).getAverage(): k(k+1)/2 + k separate mult + getAverage round trips, each a full
// device->host copy in the reference (RandomVariableCuda.java:869-883). Here every path's k+1 values are read
// once, the products are formed in float (one rounding each, like RVF:1073-1076) and summed in double.
// k <= 12, so this is a bandwidth-bound reduction (4*(k+1) bytes per path), not a tensor-core contraction.
#include <cuda_runtime.h>

#include "kernels.h"

namespace fmc {

namespace {

constexpr int RTHREADS = 256;
constexpr int RELEMS = 4;                        // one float4 per vector per thread
constexpr int RTILE = RTHREADS * RELEMS;         // 1024 paths per block iteration

template <int K>
__global__ void __launch_bounds__(RTHREADS)
regression_kernel(const __grid_constant__ RegressionParams P)
{
    constexpr int M = K * (K + 1) / 2 + K;
    double acc[M];
#pragma unroll
    for (int t = 0; t < M; t++) acc[t] = 0.0;

    const int tid = threadIdx.x;
    const long long n = P.n;
    const long long n_tiles = (n + RTILE - 1) / RTILE;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long base = tile * RTILE + (long long)tid * 4;
        float b[K][RELEMS], y[RELEMS];
        const bool full = base + 4 <= n;
#pragma unroll
        for (int i = 0; i < K; i++) {
            const float* p = P.basis[i];
            if (p == nullptr) {
#pragma unroll
                for (int e = 0; e < RELEMS; e++) b[i][e] = P.scalars[i];
            } else if (full) {
                const float4 v = *reinterpret_cast<const float4*>(p + base);
                b[i][0] = v.x; b[i][1] = v.y; b[i][2] = v.z; b[i][3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < RELEMS; e++) b[i][e] = (base + e < n) ? p[base + e] : 0.0f;
            }
        }
        if (full) {
            const float4 v = *reinterpret_cast<const float4*>(P.y + base);
            y[0] = v.x; y[1] = v.y; y[2] = v.z; y[3] = v.w;
        } else {
#pragma unroll
            for (int e = 0; e < RELEMS; e++) y[e] = (base + e < n) ? P.y[base + e] : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < RELEMS; e++) {
            if (full || base + e < n) {
                int t = 0;
#pragma unroll
                for (int i = 0; i < K; i++)
#pragma unroll
                    for (int j = i; j < K; j++, t++) acc[t] += (double)__fmul_rn(b[i][e], b[j][e]);
#pragma unroll
                for (int i = 0; i < K; i++, t++) acc[t] += (double)__fmul_rn(y[e], b[i][e]);
            }
        }
    }

    // block reduction: warp shuffle tree, then warps in order (deterministic)
    __shared__ double smem[RTHREADS / 32][M];
    __shared__ bool is_last;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int t = 0; t < M; t++) {
        double v = acc[t];
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
    // last block: entry t is summed over blocks by warp (t % 8) lanes striding the blocks, fixed order
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

int regression_max_blocks_per_sm() {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, regression_kernel<8>, RTHREADS, 0) != cudaSuccess) nb = 1;
    return nb > 0 ? nb : 1;
}

}  // namespace fmc
