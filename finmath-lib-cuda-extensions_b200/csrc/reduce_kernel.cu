// reduce_kernel.cu — streaming reduction of a MATERIALISED vector (getAverage / getVariance / getMin / getMax and the
// weighted forms on a vector that already lives in HBM). A fused "chain -> reduce" goes through the op-tape interpreter
// (tape_kernel.cu); when there is no chain to interpret, the interpreter's per-chunk dispatch is pure overhead and this
// kernel does the same job as a plain bandwidth-bound stream: 128-bit loads, 4 independent loads per thread in flight,
// double accumulation in a fixed pairwise order (deterministic for a given grid), warp shuffles, one partial per block,
// last block merges. Replaces the host loops behind RandomVariableCuda.getAverage()/getVariance()
// (/root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:869-907: full device->host copy + CPU sum)
// and the two unused shared-memory reduction kernels of RandomVariableCudaKernel.cu:268-349.
// Semantics per mode as in tape_isa.h (RM_*); result layout {count, value, M2} like the interpreter's epilogue.
//
// Bound: HBM. Algorithmic bytes: 4 per element (8 for the weighted modes).
#include <cuda_runtime.h>
#include <algorithm>
#include <math.h>

#include "kernels.h"
#include "tape_isa.h"
#include "reduce_common.cuh"

namespace fmc {

namespace {

constexpr int RT = 256;            // threads per block
constexpr int RU = 4;              // float4 loads per thread per iteration
constexpr int RTILE = RT * RU * 4; // elements per block iteration

__device__ __forceinline__ float jminf(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float jmaxf(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

__device__ Part block_reduce(int mode, Part p, Part* smem) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) p = merge(mode, p, shfl_down(p, d));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = p;
    __syncthreads();
    if (warp == 0) {
        Part q = (lane < RT / 32) ? smem[lane] : Part{0.0, 0.0, 0.0};
#pragma unroll
        for (int d = RT / 64; d > 0; d >>= 1) q = merge(mode, q, shfl_down(q, d));
        p = q;
    }
    return p;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// MODE is a compile-time ReduceMode: each variant carries only its own state
template <int MODE>
__global__ void __launch_bounds__(RT)
reduce_kernel(const __grid_constant__ ReduceParams P)
{
    const long long n = P.n;
    const int tid = threadIdx.x;
    double acc = 0.0, s1 = 0.0, s2 = 0.0, shiftK = 0.0;
    float fext = 0.0f;
    long long cnt = 0;
    constexpr bool W = (MODE == RM_DOT || MODE == RM_WSQ);

    auto fold4 = [&](const float4 x, const float4 w) {
        if (MODE == RM_SUM) acc += ((double)x.x + (double)x.y) + ((double)x.z + (double)x.w);
        else if (MODE == RM_MOMENTS) {
            if (cnt == 0) shiftK = (double)x.x;
            const double d0 = (double)x.x - shiftK, d1 = (double)x.y - shiftK, d2 = (double)x.z - shiftK, d3 = (double)x.w - shiftK;
            s1 += (d0 + d1) + (d2 + d3);
            s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        } else if (MODE == RM_MIN) { const float m = jminf(jminf(x.x, x.y), jminf(x.z, x.w)); fext = cnt == 0 ? m : jminf(fext, m); }
        else if (MODE == RM_MAX) { const float m = jmaxf(jmaxf(x.x, x.y), jmaxf(x.z, x.w)); fext = cnt == 0 ? m : jmaxf(fext, m); }
        else if (MODE == RM_DOT) acc += ((double)x.x * (double)w.x + (double)x.y * (double)w.y) + ((double)x.z * (double)w.z + (double)x.w * (double)w.w);
        else {
            const double d0 = (double)x.x - P.param, d1 = (double)x.y - P.param, d2 = (double)x.z - P.param, d3 = (double)x.w - P.param;
            acc += (d0 * d0 * (double)w.x + d1 * d1 * (double)w.y) + (d2 * d2 * (double)w.z + d3 * d3 * (double)w.w);
        }
        cnt += 4;
    };
    auto fold1 = [&](const float x, const float w) {
        if (MODE == RM_SUM) acc += (double)x;
        else if (MODE == RM_MOMENTS) { if (cnt == 0) shiftK = (double)x; const double d = (double)x - shiftK; s1 += d; s2 += d * d; }
        else if (MODE == RM_MIN) fext = cnt == 0 ? x : jminf(fext, x);
        else if (MODE == RM_MAX) fext = cnt == 0 ? x : jmaxf(fext, x);
        else if (MODE == RM_DOT) acc += (double)x * (double)w;
        else { const double d = (double)x - P.param; acc += d * d * (double)w; }
        cnt += 1;
    };

    const long long n_tiles = n / RTILE;            // full tiles: 4 independent 128-bit loads per thread (8 when weighted)
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long base = t * RTILE + (long long)tid * 4;
        float4 x[RU], w[RU];
#pragma unroll
        for (int u = 0; u < RU; u++) {
            x[u] = ldg4(P.x + base + (long long)u * RT * 4);
            w[u] = W ? ldg4(P.w + base + (long long)u * RT * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < RU; u++) fold4(x[u], w[u]);
    }
    // ragged tail (less than one tile): block 0 walks it element by element
    if (blockIdx.x == 0)
        for (long long i = n_tiles * RTILE + tid; i < n; i += RT) fold1(P.x[i], W ? P.w[i] : 0.f);

    Part part;
    part.c = (double)cnt; part.v = acc; part.m = 0.0;
    if (MODE == RM_MIN || MODE == RM_MAX) part.v = (double)fext;
    if (MODE == RM_MOMENTS && cnt > 0) { part.v = shiftK + s1 / part.c; part.m = s2 - s1 * s1 / part.c; }
    constexpr int MM = W ? RM_SUM : MODE;

    __shared__ Part red_smem[RT / 32];
    __shared__ bool is_last;
    Part blk = block_reduce(MM, part, red_smem);
    if (tid == 0) {
        double* dst = P.partials + 4ll * blockIdx.x;
        dst[0] = blk.c; dst[1] = blk.v; dst[2] = blk.m;
        __threadfence();
        const unsigned ticket = atomicAdd(P.counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Part q = {0.0, 0.0, 0.0};
    for (unsigned k = tid; k < gridDim.x; k += RT) {
        const volatile double* src = P.partials + 4ll * k;
        Part t = { src[0], src[1], src[2] };
        q = merge(MM, q, t);
    }
    q = block_reduce(MM, q, red_smem);
    if (tid == 0) {
        *P.counter = 0u;
        finish_reduction(MM, q, P.xchg, P.ticket, P.result, P.host_result);
    }
}

// ---- sums of K vectors in one launch (kernels.h: BatchSumParams) ----
__global__ void __launch_bounds__(RT)
batch_sum_kernel(const __grid_constant__ BatchSumParams P)
{
    const int j = blockIdx.x / P.blocks_per_vec, b = blockIdx.x - j * P.blocks_per_vec, B = P.blocks_per_vec;
    const float* __restrict__ x = P.x[j];
    const long long n = P.n;
    const int tid = threadIdx.x;
    double acc = 0.0;
    long long cnt = 0;                                                  // merge() skips empty partials: keep the counts like reduce_kernel
    const long long n_tiles = n / RTILE;
    for (long long t = b; t < n_tiles; t += B) {                       // the walk of reduce_kernel<RM_SUM> with gridDim.x == B
        const long long base = t * RTILE + (long long)tid * 4;
        float4 v[RU];
#pragma unroll
        for (int u = 0; u < RU; u++) v[u] = ldg4(x + base + (long long)u * RT * 4);
#pragma unroll
        for (int u = 0; u < RU; u++) acc += ((double)v[u].x + (double)v[u].y) + ((double)v[u].z + (double)v[u].w);
        cnt += 4 * RU;
    }
    if (b == 0)
        for (long long i = n_tiles * RTILE + tid; i < n; i += RT) { acc += (double)x[i]; cnt++; }

    __shared__ Part red_smem[RT / 32];
    __shared__ bool is_last;
    Part part = {(double)cnt, acc, 0.0};
    Part blk = block_reduce(RM_SUM, part, red_smem);
    if (tid == 0) {
        double* dst = P.partials + 2 * ((long long)j * B + b);
        dst[0] = blk.c; dst[1] = blk.v;
        __threadfence();
        is_last = (atomicAdd(P.counters + j, 1u) == (unsigned)B - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Part q = {0.0, 0.0, 0.0};
    for (int k = tid; k < B; k += RT) {
        const volatile double* src = P.partials + 2 * ((long long)j * B + k);
        Part t = {src[0], src[1], 0.0};
        q = merge(RM_SUM, q, t);
    }
    q = block_reduce(RM_SUM, q, red_smem);
    if (tid == 0) {
        P.counters[j] = 0u;
        volatile double* h = P.host_out;
        h[j] = q.v;
        __threadfence_system();
        if (atomicAdd(P.counters + BATCH_MAX, 1u) == (unsigned)P.k - 1u) {
            __threadfence_system();
            P.counters[BATCH_MAX] = 0u;
            h[BATCH_MAX] = P.ticket;
        }
    }
}

}  // namespace

cudaError_t launch_batch_sum(const BatchSumParams& P, cudaStream_t stream) {
    if (P.k < 1 || P.k > BATCH_MAX || P.blocks_per_vec < 1) return cudaErrorInvalidValue;
    batch_sum_kernel<<<P.k * P.blocks_per_vec, RT, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_reduce(const ReduceParams& P, int grid, cudaStream_t stream) {
    switch (P.mode) {
    case RM_SUM: reduce_kernel<RM_SUM><<<grid, RT, 0, stream>>>(P); break;
    case RM_MOMENTS: reduce_kernel<RM_MOMENTS><<<grid, RT, 0, stream>>>(P); break;
    case RM_MIN: reduce_kernel<RM_MIN><<<grid, RT, 0, stream>>>(P); break;
    case RM_MAX: reduce_kernel<RM_MAX><<<grid, RT, 0, stream>>>(P); break;
    case RM_DOT: reduce_kernel<RM_DOT><<<grid, RT, 0, stream>>>(P); break;
    case RM_WSQ: reduce_kernel<RM_WSQ><<<grid, RT, 0, stream>>>(P); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
int reduce_tile_elems() { return RTILE; }

// ---- (float) cast of an uploaded double chunk (asynchronous pinned upload path, runtime.cpp: upload_pinned) ----
// cvt.rn.f32.f64 rounds to nearest even like Java's (float) cast (RandomVariableCuda.java:768-774).
namespace {
__global__ void __launch_bounds__(256) cast_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 2;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
        if (i + 1 < n) {
            const double2 v = *reinterpret_cast<const double2*>(src + i);
            *reinterpret_cast<float2*>(dst + i) = make_float2(__double2float_rn(v.x), __double2float_rn(v.y));
        } else dst[i] = __double2float_rn(src[i]);
    }
}
}  // namespace

cudaError_t launch_cast_f64_f32(const double* src, float* dst, long long n, int sm_count, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const long long blocks = (n / 2 + 255) / 256;
    const int grid = (int)std::max<long long>(1, std::min<long long>(blocks, (long long)sm_count * 8));
    cast_f64_f32_kernel<<<grid, 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}

}  // namespace fmc
