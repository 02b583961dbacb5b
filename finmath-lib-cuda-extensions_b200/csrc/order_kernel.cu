// order_kernel.cu — order statistics on the device: getQuantile / getQuantileExpectation / getHistogram without the
// full device->host copy and the O(n log n) host sort of the reference
// (/root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:970-1091; specification
// RandomVariableFromFloatArray.java:472-602: Arrays.sort order, i.e. -0.0f < +0.0f and NaN last).
//
// Floats are mapped to unsigned keys with the same total order as java.util.Arrays.sort(float[]) (sign flip; every NaN
// becomes the largest key). The element of rank r is found by a 4-pass most-significant-digit radix SELECT: each pass
// streams the vector once, histograms one 8-bit digit of the keys that match the digits fixed so far (shared-memory
// histogram per block, one atomic add per non-empty bin per block), and the host — or, for a vector sharded over several
// GPUs, the host after an all-reduce of the 256 counts — picks the bin that contains the rank. Counts are integers, so
// the result is exactly the element a sort would put there. The quantile expectation needs two selects plus one pass
// that sums what lies strictly between the two values; the histogram is one pass with a binary search per element.
//
// Bound: HBM, 4 bytes per element and pass.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace fmc {

namespace {

constexpr int OT = 256;

__device__ __forceinline__ uint32_t sort_key(float f) {
    if (f != f) return 0xffffffffu;
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(OT)
select_hist_kernel(const float* __restrict__ x, long long n, uint32_t prefix, uint32_t mask, int shift, double* __restrict__ hist /* [256] */)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n4; i += (long long)gridDim.x * OT) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t key = sort_key(e[k]);
            if (((key ^ prefix) & mask) == 0u) atomicAdd(&h[(key >> shift) & 255u], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < n; i += OT) {
            const uint32_t key = sort_key(x[i]);
            if (((key ^ prefix) & mask) == 0u) atomicAdd(&h[(key >> shift) & 255u], 1u);
        }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (double)h[threadIdx.x]);     // integer-valued doubles: exact, order free
}

// out[0] = #{key < lo}, out[1] = #{key == lo}, out[2] = #{lo < key < hi}, out[3] = sum of x with lo < key < hi, out[4] = #{key == hi}
__global__ void __launch_bounds__(OT)
range_stats_kernel(const float* __restrict__ x, long long n, uint32_t lo, uint32_t hi, double* __restrict__ out)
{
    double c_lt = 0, c_lo = 0, c_mid = 0, s_mid = 0, c_hi = 0;
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n; i += (long long)gridDim.x * OT) {
        const float v = x[i];
        const uint32_t key = sort_key(v);
        if (key < lo) c_lt += 1.0;
        else if (key == lo) c_lo += 1.0;
        else if (key < hi) { c_mid += 1.0; s_mid += (double)v; }
        else if (key == hi) c_hi += 1.0;
    }
    double vals[5] = {c_lt, c_lo, c_mid, s_mid, c_hi};
    __shared__ double sm[5][OT / 32];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double v = vals[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < OT / 32; w++) v += sm[threadIdx.x][w];
        atomicAdd(&out[threadIdx.x], v);            // counts exact; the sum's rounding depends on block order (1e-16 relative)
    }
}

// counts[k] = #{pts[k-1] < x <= pts[k]} (k = 0: x <= pts[0]), counts[m] = the rest (x > pts[m-1] or NaN); RVF:558-569
__global__ void __launch_bounds__(OT)
histogram_kernel(const float* __restrict__ x, long long n, const double* __restrict__ pts, int m, double* __restrict__ counts)
{
    extern __shared__ unsigned char smem_raw[];
    double* sp = reinterpret_cast<double*>(smem_raw);
    unsigned int* sc = reinterpret_cast<unsigned int*>(sp + m);
    for (int k = threadIdx.x; k < m; k += OT) sp[k] = pts[k];
    for (int k = threadIdx.x; k <= m; k += OT) sc[k] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n; i += (long long)gridDim.x * OT) {
        const double v = (double)x[i];
        int lo = 0, hi = m;                         // first k with v <= pts[k]; m if none (also for NaN)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (v <= sp[mid]) hi = mid; else lo = mid + 1; }
        atomicAdd(&sc[lo], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= m; k += OT) if (sc[k]) atomicAdd(&counts[k], (double)sc[k]);
}

}  // namespace

cudaError_t launch_select_hist(const float* x, long long n, uint32_t prefix, uint32_t mask, int shift, double* hist, int grid, cudaStream_t s) {
    select_hist_kernel<<<grid, OT, 0, s>>>(x, n, prefix, mask, shift, hist);
    return cudaGetLastError();
}
cudaError_t launch_range_stats(const float* x, long long n, uint32_t lo, uint32_t hi, double* out, int grid, cudaStream_t s) {
    range_stats_kernel<<<grid, OT, 0, s>>>(x, n, lo, hi, out);
    return cudaGetLastError();
}
cudaError_t launch_histogram(const float* x, long long n, const double* pts, int m, double* counts, int grid, cudaStream_t s) {
    histogram_kernel<<<grid, OT, sizeof(double) * (size_t)m + sizeof(unsigned int) * (size_t)(m + 1), s>>>(x, n, pts, m, counts);
    return cudaGetLastError();
}

}  // namespace fmc
