// kernels.h — host-callable launchers of the sm_100a kernels (internal to libfmcuda.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tape_isa.h"

namespace fmc {

// tape_launch.cu + tape_kernel_e16/8/4.cu (one body: tape_interp.cuh)
// n_warps per CTA: 1..TAPE_MAX_WARPS; geometry: P.elems; a tape too long for the inline argument block is uploaded on copy_stream (nullptr: on stream)
cudaError_t launch_tape(const TapeParams& P, int grid, int n_warps, cudaStream_t stream, cudaStream_t copy_stream);
cudaError_t tape_kernel_setup(size_t* max_smem_per_cta);   // opts the kernels in to the device's full shared memory, allocates the tape ring
void tape_kernel_teardown();
#ifdef FMC_TAPE_TIMING
cudaError_t tape_read_stamps_e16(unsigned long long* out);   // development aid (Makefile: timing): globaltimer stamps of the last fused reduction
#endif
size_t tape_smem_bytes(int n_ptrs, int n_instr, int n_slots, int n_sets, int n_warps, int elems);   // dynamic shared memory of one CTA
int tape_max_blocks_per_sm(size_t smem_bytes, int reduce_mode, int n_warps, int elems);

// reduce_kernel.cu — streaming reduction of a materialised vector (no chain to interpret)
struct ReduceParams {
    long long n;
    int mode;                 // ReduceMode
    double param;             // RM_WSQ: the mean
    const float* x;
    const float* w;           // weights (RM_DOT / RM_WSQ) or nullptr
    double* partials;         // [gridDim.x][4]
    unsigned int* counter;
    double* result;           // {count, value, M2}
    double* host_result;      // mapped pinned mirror (+ ticket in [3]) or nullptr
    double ticket;
    Exchange xchg;            // peer tables of a path-sharded run (reduce_common.cuh)
};
cudaError_t launch_reduce(const ReduceParams& P, int grid, cudaStream_t stream);
// The sums of K materialised vectors of the same length in ONE launch (Runtime::reduce_batch: the averages of all vectors one
// flush produced, e.g. the calibration products of one objective-function evaluation). blocks_per_vec CTAs work on each vector
// exactly as launch_reduce's grid of that size would (same tiles, same order of additions: the sum of a vector does not
// depend on whether it was reduced alone or in a batch). The last CTA of each vector writes its sum to host_out[j] (mapped
// pinned memory), the last of those writes the ticket to host_out[BATCH_MAX].
constexpr int BATCH_MAX = 256;
struct BatchSumParams {
    long long n;
    int k;                      // vectors, <= BATCH_MAX
    int blocks_per_vec;
    double* partials;           // [k][blocks_per_vec][2] {count, sum}
    unsigned int* counters;     // [BATCH_MAX + 1], zero between launches
    double* host_out;           // [BATCH_MAX + 1] mapped pinned memory: sums, then the ticket
    double ticket;
    const float* x[BATCH_MAX];
};
cudaError_t launch_batch_sum(const BatchSumParams& P, cudaStream_t stream);
// dst[i] = (float)src[i]; both device pointers, 8-byte aligned (dst offsets of the upload path are multiples of the chunk size)
cudaError_t launch_cast_f64_f32(const double* src, float* dst, long long n, int sm_count, cudaStream_t stream);
int reduce_tile_elems();

// order_kernel.cu — radix select / range statistics / histogram (order statistics without a sort)
cudaError_t launch_select_hist(const float* x, long long n, uint32_t prefix, uint32_t mask, int shift, double* hist /* [256], zeroed */, int grid, cudaStream_t s);
cudaError_t launch_range_stats(const float* x, long long n, uint32_t key_lo, uint32_t key_hi, double* out /* [5], zeroed */, int grid, cudaStream_t s);
cudaError_t launch_histogram(const float* x, long long n, const double* pts /* device [m] */, int m, double* counts /* [m+1], zeroed */, int grid, cudaStream_t s);

// regression_kernel.cu — fused normal equations: one pass over k basis vectors + y.
constexpr int REG_MAX_K = 12;
struct RegressionParams {
    long long n;
    int k;
    const float* basis[REG_MAX_K];   // nullptr -> deterministic scalar
    float scalars[REG_MAX_K];
    const float* y;
    double* partials;                // [gridDim.x][REG_MAX_K*(REG_MAX_K+1)/2 + REG_MAX_K]
    unsigned int* counter;
    double* result;                  // [k*(k+1)/2 + k] sums (not yet divided by n)
    int float_products;              // != 0: sums of the float products fl32(a*b) (RandomVariableFromFloatArray); 0: of the exact products
};
cudaError_t launch_regression(const RegressionParams& P, int grid, cudaStream_t stream);
int regression_max_blocks_per_sm(int k);
int regression_tile_elems();

// brownian_kernel.cu — MT19937 (commons-math3 stream) with jump-ahead + AS241 inverse normal.
constexpr int MT_N = 624;
constexpr int MT_CHUNK_WORDS = 1 << 15;          // jump granularity: block start states sit on multiples of this
struct BrownianParams {
    const uint32_t* block_states;   // [n_blocks][624] generator state (window x[cW .. cW+623]) at each block's start chunk
    const long long* chunk_of_block;// [n_blocks] start chunk index c of each block
    int n_blocks;
    long long paths_per_block;
    long long p0, np;               // first path and number of paths of the slice this process owns
    int T, F;
    int PT;                         // paths per shared-memory tile (odd; T*F*PT >= 312)
    const double* sqrt_dt;          // [T] device
    float* const* out;              // [T*F] device array of device pointers, each np floats
};
cudaError_t launch_brownian(const BrownianParams& P, cudaStream_t stream);
int brownian_max_blocks_per_sm(int T, int F, int PT);     // resident blocks per SM for this tile geometry (occupancy query)
// out_states[b] = seeded state advanced by chunk_of_block[b] * MT_CHUNK_WORDS words
cudaError_t launch_mt_jump(const uint32_t* base_state, const uint32_t* polys, int npoly, const long long* chunk_of_block,
                           uint32_t* out_states, int n_blocks, cudaStream_t stream);
// tempered stream words [skip, skip+count); block b emits words_per_block words starting at skip + b*words_per_block
cudaError_t launch_mt_raw(const uint32_t* block_states, const long long* chunk_of_block, int n_blocks, long long words_per_block,
                          unsigned long long skip, long long count, uint32_t* out, cudaStream_t stream);

}  // namespace fmc
