/*
 * fmcuda.h — C ABI of the B200-native vector runtime behind finmath-lib's RandomVariable / BrownianMotion layer.
 *
 * This header is the drop-in boundary: it replaces everything the reference reaches through JCuda/JCurand JNI
 * (cuInit/cuCtxCreate/cuModuleLoad/cuMemAlloc/cuMemcpyHtoD/DtoH/cuLaunchKernel/curandGenerateNormal) from
 *   RVC = /root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java
 *   BMC = /root/reference/src/main/java/net/finmath/cuda/montecarlo/alternative/BrownianMotionCudaWithRandomVariableCuda.java
 * A Java shim (JNI or java.lang.foreign) binds exactly these symbols; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns an int status (FMC_OK == 0, negative == error) and
 *     fmc_last_error() returns a thread-local message for the last failing call of the calling thread.
 *   - fmc_vec is an opaque 64-bit handle to an IMMUTABLE fp32 device vector (the "stochastic" RandomVariable of
 *     RVC:572-574). Deterministic random variables never reach this API: as in RVC:576-577 they stay doubles on
 *     the caller side and enter as the `double s` arguments, cast to float inside (RVC:521 / RVF:789).
 *   - handles are reference counted (retain/release); every operation returns a NEW handle with refcount 1.
 *   - operations are RECORDED, not executed: they append to an op-tape (a DAG of pending nodes). The tape is
 *     executed by one fused interpreter kernel when a value is demanded (reduction, host read, device-pointer
 *     export, explicit fmc_flush) or when the number of pending nodes exceeds the "flush_threshold" option.
 *   - all entry points are thread safe and callable from any thread (one runtime lock; the calling thread is bound to
 *     the runtime's device); a reduction waits for its result WITHOUT the lock in single-rank runs, so other threads
 *     keep recording and launching meanwhile. Host buffers are only read/written during the call.
 *   - arithmetic contract: IEEE binary32 round-to-nearest, NO fused multiply-add (the reference compiles with
 *     `-fmad false`, JCudaUtils.java:65-75, to match Java float arithmetic); exp/log/pow/sin/cos are evaluated in
 *     double and rounded to float as RandomVariableFromFloatArray.java:849,890,905,920,935,950 does.
 */
#ifndef FMCUDA_H
#define FMCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t fmc_vec;              /* 0 is never a valid handle */

/* ---- status codes ---- */
#define FMC_OK               0
#define FMC_ERR_INVALID     (-1)       /* bad handle / argument                                  -> IllegalArgumentException */
#define FMC_ERR_OOM         (-2)       /* device allocation failed (RVC:358-376)                  -> OutOfMemoryError */
#define FMC_ERR_CUDA        (-3)       /* CUDA runtime error (JCuda CudaException, RVC:167)        -> RuntimeException */
#define FMC_ERR_SIZE        (-4)       /* operand size mismatch                                   -> ArrayIndexOutOfBoundsException */
#define FMC_ERR_NOT_INIT    (-5)       /* fmc_init not called / no CUDA device: there is NO CPU fallback */
#define FMC_ERR_COMM        (-6)       /* NCCL error */
#define FMC_ERR_UNSUPPORTED (-7)

/* ---- opcodes (same numbering as oracle/fm_oracle.h) ---- */
enum {
    /* fmc_op_vs: vector (op) scalar.  RVC:1172-1277 -> RandomVariableCudaKernel.cu:2-106 */
    FMC_CAP = 1, FMC_FLOOR = 2, FMC_ADD = 3, FMC_SUB = 4, FMC_BUS = 5, FMC_MULT = 6, FMC_DIV = 7, FMC_VID = 8, FMC_POW = 9,
    /* fmc_op_v: unary.  RVC:1285-1352 -> kernel.cu:109-156; sin/cos/isNaN exist only in RVF:927-954,1440-1451 */
    FMC_SQUARED = 20, FMC_SQRT = 21, FMC_EXP = 22, FMC_LOG = 23, FMC_SIN = 24, FMC_COS = 25, FMC_INVERT = 26, FMC_ABS = 27, FMC_ISNAN = 28,
    /* fmc_op_vvs / fmc_op_vvv.  RVC:1583-1695 -> kernel.cu:224-284; choose: RVF:1264-1285 (RVC:1632 returns null) */
    FMC_ACCRUE = 40, FMC_DISCOUNT = 41, FMC_ADDPRODUCT = 42, FMC_CHOOSE = 43, FMC_ADDRATIO = 44, FMC_SUBRATIO = 45
};

/* ---- reduction kinds for fmc_reduce ---- */
enum {
    FMC_RED_SUM = 1,             /* sum of elements (double)                                              */
    FMC_RED_AVERAGE = 2,         /* getAverage()           RVC:869-883 -> RVF:314-334; NaN if n == 0       */
    FMC_RED_VARIANCE = 3,        /* getVariance()          RVF:360-382 (biased, /n); 0 if n == 1           */
    FMC_RED_SAMPLE_VARIANCE = 4, /* getSampleVariance()    RVF:410-419                                     */
    FMC_RED_MIN = 5,             /* getMin()               RVF:284-296                                     */
    FMC_RED_MAX = 6,             /* getMax()               RVF:299-311                                     */
    FMC_RED_AVERAGE_W = 7,       /* getAverage(prob)       RVF:337-357: sum(x*p)/n                         */
    FMC_RED_VARIANCE_W = 8       /* getVariance(prob)      RVF:385-407: sum((x-avg_w)^2*p)  (not / n)      */
};

/* ---- lifecycle (replaces RVC.DeviceMemoryPool ctor RVC:160-264 and the shutdown hook RVC:252-261) ---- */
/* device_index < 0 counts from the end (-1 = last device), as the system property
 * net.finmath.montecarlo.opencl.RandomVariableCuda.deviceIndex does (RVC:161,177). Idempotent. */
int fmc_init(int device_index);
int fmc_shutdown(void);
int fmc_is_initialized(void);
const char* fmc_last_error(void);
int fmc_device_count(int* count);
int fmc_device_info(char* name, size_t name_len, int* sm_count, uint64_t* total_mem_bytes, int* cc_major, int* cc_minor);

/* ---- vectors (replaces RVC:618-734 constructors, RVC:457-481 H2D/D2H, RVC:737-749 getDevicePointer) ---- */
int fmc_vec_from_f64(const double* host, int64_t n, fmc_vec* out);   /* (float) cast RVC:768-774, H2D via pinned staging */
/* Pinned host memory and the asynchronous upload from it. fmc_vec_from_f64 reads the caller's (pageable) array during the call:
 * the (float) cast of RVC:768-774 runs on host threads into pinned staging, which bounds an upload-heavy run by host memory
 * bandwidth. A caller that keeps its doubles in memory from fmc_host_alloc can use fmc_vec_from_f64_pinned instead: the doubles
 * are copied by DMA on the copy stream and cast on the device, overlapping the kernels already queued. THE BUFFER IS READ AFTER
 * THE CALL RETURNS: keep it unchanged until a reduction / host read that depends on the vector has returned, or fmc_sync(). */
int fmc_host_alloc(size_t bytes, void** out);
int fmc_host_free(void* p);
int fmc_vec_from_f64_pinned(const double* pinned_host, int64_t n, fmc_vec* out);
int fmc_vec_from_f32(const float* host, int64_t n, fmc_vec* out);
int fmc_vec_fill(double value, int64_t n, fmc_vec* out);             /* RVF:139-146 (numberOfPath, value) constructor */
int fmc_vec_alloc(int64_t n, fmc_vec* out);                          /* RVC:737-739 getDevicePointer(size): uninitialised */
int fmc_vec_retain(fmc_vec v);
int fmc_vec_release(fmc_vec v);                                      /* what the pool's ReferenceQueue does in RVC:295-306 */
int fmc_vec_size(fmc_vec v, int64_t* n);
int fmc_vec_to_f64(fmc_vec v, double* host, int64_t n);              /* getRealizations() RVC:1115-1122 (flushes) */
int fmc_vec_to_f32(fmc_vec v, float* host, int64_t n);               /* getValuesAsFloat RVC:469-481 */
int fmc_vec_get(fmc_vec v, int64_t i, double* out);                  /* get(i) RVF:266-272 (RVC:812-818 throws) */
int fmc_vec_device_ptr(fmc_vec v, void** device_ptr);                /* materialises; pointer valid while v is retained */

/* ---- recorded elementwise operations ---- */
int fmc_op_vs (int opcode, fmc_vec a, double s, fmc_vec* out);
int fmc_op_v  (int opcode, fmc_vec a, fmc_vec* out);
int fmc_op_vv (int opcode, fmc_vec a, fmc_vec b, fmc_vec* out);                 /* add sub bus mult div vid cap floor */
int fmc_op_vvs(int opcode, fmc_vec a, fmc_vec b, double s, fmc_vec* out);       /* accrue discount addProduct(v, scalar) */
int fmc_op_vvv(int opcode, fmc_vec a, fmc_vec b, fmc_vec c, fmc_vec* out);      /* addProduct choose addRatio subRatio */
/* choose with deterministic branches: a handle of 0 selects the scalar (RVF:1281 reads get(i) of either kind) */
int fmc_op_choose(fmc_vec trigger, fmc_vec if_nonneg, double s_nonneg, fmc_vec if_neg, double s_neg, fmc_vec* out);

/* ---- reductions (flush the cone of `a`; the last op chain is fused into the reduction kernel) ---- */
int fmc_reduce(int kind, fmc_vec a, fmc_vec weights_or_0, double* out);
/* order statistics (RVF:472-602); device sort */
int fmc_quantile(fmc_vec a, double q, double* out);
int fmc_quantile_expectation(fmc_vec a, double q0, double q1, double* out);
int fmc_histogram(fmc_vec a, const double* interval_points, int m, double* out /* m+1 */);

/* ---- conditional expectation regression: normal equations in one fused pass ----
 * XtX[i*k+j] = average(b_i*b_j), Xty[i] = average(y*b_i) with float products and double sums
 * (finmath-lib MonteCarloConditionalExpectationRegression; hook RVF:861-864). basis[i]==0 -> constant scalars[i]. */
int fmc_regression_normal_eq(const fmc_vec* basis, const double* scalars, int k, fmc_vec y, double* XtX, double* Xty);

/* ---- Brownian increments: on-device MT19937 (commons-math3 stream) + AS241 inverse normal ----
 * replaces BMC:141-182 (cuRAND XORWOW) by the stream of finmath-lib BrownianMotionFromMersenneRandomNumbers
 * (call sites LIBORMarketModelCalibrationATMTest.java:283, MonteCarloBlackScholesModelTest.java:78-85).
 * seed_mode 0: MersenneTwister(long) == init_by_array{hi,lo};  1: MersenneTwister(int) == init_genrand.
 * Generates the increments of paths [p0,p1) of an n_total-path motion: out[t*F+f] has p1-p0 elements. */
int fmc_brownian_generate(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1,
                          const double* sqrt_dt /* T */, fmc_vec* out /* T*F */);
/* raw tempered uint32 words [skip, skip+count) of the same generator (bit-exactness witness) */
int fmc_mt19937_raw(int seed_mode, int64_t seed, uint64_t skip, int64_t count, uint32_t* host_out);

/* ---- execution control ---- */
int fmc_flush(void);                    /* execute every pending node that is still referenced */
int fmc_sync(void);                     /* flush + wait for the device (cuCtxSynchronize, RVC:472-476) */
/* options: "flush_threshold" (pending nodes before an automatic flush; default 4096),
 *          "fuse" (1 default; 0 = execute every op as its own kernel, the reference's execution model),
 *          "profile" (0 default; see fmc_profile_read);
 *          interpreter scheduling knobs (tuning / tests; defaults in csrc/runtime.h): "ring_max", "ring_min", "target_ctas",
 *          "horizon", "pipeline", "max_sets", "max_regs", "grid_limit", "fuse_ops", "cta_warps", "zero_copy_reduce", "leaf_reduce_kernel",
 *          "p2p_reduce" (1 default: sharded runs exchange reduction partials inside the kernel over NVLink peer memory; 0: NCCL);
 *          "tape_cache" (1 default: a cone of pending nodes whose structure was lowered before is replayed with the new
 *          buffers and immediates patched in instead of being code-generated again);
 *          read-only: "p2p_ready", "device_index", "tape_cache_hits", "tape_cache_misses", "tape_cache_entries";
 *          read-only host-side timers in microseconds since fmc_reset_stats: "host_us_codegen", "host_us_launch", "host_us_sync",
 *          "host_us_upload" (host->device calls; of that "host_us_upload_wait" waiting for a free pinned staging chunk).
 *          The environment variable FMC_OPTIONS="key=value,..." is applied once at fmc_init. */
int fmc_set_option(const char* key, double value);
int fmc_get_option(const char* key, double* value);

typedef struct fmc_stats {
    uint64_t bytes_in_use;              /* device bytes held by live vectors */
    uint64_t bytes_cached;              /* device bytes in the pool's free lists */
    uint64_t bytes_reserved;            /* device bytes obtained from the driver (slabs) */
    uint64_t bytes_high_water;
    uint64_t n_alloc, n_alloc_reused;   /* pool requests / served from a free list */
    uint64_t n_ops_recorded;            /* elementwise ops recorded */
    uint64_t n_kernels;                 /* kernels launched (all kinds) */
    uint64_t n_tape_kernels;            /* interpreter launches */
    uint64_t n_tape_instr;              /* interpreter instructions issued (sum over launches) */
    uint64_t n_nodes_stored;            /* nodes written to HBM */
    uint64_t n_nodes_fused;             /* nodes that lived only in registers */
    uint64_t n_flushes;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t live_handles, pending_nodes;
} fmc_stats;
int fmc_get_stats(fmc_stats* out);
int fmc_reset_stats(void);
int fmc_pool_trim(void);                /* RandomVariableCuda.clean() RVC:751-753: return cached blocks to the driver */
int fmc_pool_purge(void);               /* RandomVariableCuda.purge() RVC:755-757 */

/* option "profile" = 1: every interpreter launch is bracketed by CUDA events; fmc_profile_read synchronises and returns
 * the summed kernel time, the algorithmic bytes (4 * n * (leaf vectors read + result vectors stored)) and the launch
 * count since the last read, then resets the counters. */
int fmc_profile_read(double* tape_ms, uint64_t* tape_algorithmic_bytes, uint64_t* tape_launches);

/* device-side timing of the compute stream (CUDA events), for benchmarks */
int fmc_timer_start(void);
int fmc_timer_stop(float* elapsed_ms);  /* synchronises on the stop event */

/* ---- multi GPU: one process per GPU, each holding a contiguous path slice of every vector ----
 * After fmc_comm_init, fmc_reduce gathers the ranks' {count, value, M2} partials (one ncclAllGather on the compute
 * stream, merged in rank order) and fmc_regression_normal_eq all-reduces its sums (ncclAllReduce); both return the
 * statistics of the GLOBAL vector. Vector data never crosses NVLink. */
#define FMC_UNIQUE_ID_BYTES 128
int fmc_comm_get_unique_id(char* id /* FMC_UNIQUE_ID_BYTES */);
int fmc_comm_init(int rank, int nranks, const char* id);
int fmc_comm_destroy(void);
int fmc_comm_info(int* rank, int* nranks);

#ifdef __cplusplus
}
#endif
#endif /* FMCUDA_H */
