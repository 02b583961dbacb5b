// runtime.h — internal structures of libfmcuda.so: device runtime (pool, stream, staging), the pending-op
// graph (op-tape) and its code generator. Replaces RandomVariableCuda.DeviceMemoryPool
// (/root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:119-558).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <cstring>
#include <deque>
#include <vector>

#include "../../include/fmcuda.h"
#include "kernels.h"
#include "tape_isa.h"

namespace fmc {

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();
struct Fail { int code; };   // thrown internally, converted to a status code at the C boundary
[[noreturn]] void fail(int code, const char* fmt, ...);
#define FMC_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    ::fmc::fail(e_ == cudaErrorMemoryAllocation ? FMC_ERR_OOM : FMC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ---------------------------------------------------------------------------------------------------------
// device memory pool: slabs obtained from the driver, carved into blocks; freed blocks are recycled by exact
// rounded size (Monte-Carlo vectors of one simulation all have the same size, like RVC:295-306) and are safe to
// reuse immediately because every kernel and copy runs in order on the one compute stream (stream-ordered
// reuse, the property RVC relies on implicitly). No cudaMemGetInfo per allocation (RVC:308), no GC coupling.
// ---------------------------------------------------------------------------------------------------------
class DevicePool {
public:
    void* alloc(size_t bytes);          // throws Fail{FMC_ERR_OOM}
    void  free(void* p);
    // For a buffer that a SECOND stream will write (host uploads on the copy stream): a cached block is only handed out if
    // it was freed at or before `settled_stamp` (every kernel that touched it is known to have finished), else fresh memory
    // is carved. *settled = false when neither was possible and an ordinary cached block is returned: the caller must then
    // order the second stream behind the compute stream.
    void* alloc_settled(size_t bytes, uint64_t settled_stamp, bool* settled);
    uint64_t free_stamp = 0;            // every free() gets the next stamp
    void  trim();                       // return slabs without live blocks to the driver
    void  purge();                      // trim, must be called with no live blocks to release everything
    uint64_t bytes_in_use = 0, bytes_cached = 0, bytes_reserved = 0, high_water = 0, n_alloc = 0, n_reused = 0;
private:
    struct Slab { char* base; size_t size, used; int live; bool dedicated; };
    struct Block { size_t size; int slab; };
    std::vector<Slab> slabs_;
    std::unordered_map<void*, Block> blocks_;                  // live + cached blocks by address
    struct Cached { void* p; int slab; uint64_t stamp; };     // what alloc() needs, so that reuse costs no map lookup
    std::unordered_map<size_t, std::vector<Cached>> free_;     // rounded size -> cached blocks (oldest first)
    size_t last_size_ = 0; std::vector<Cached>* last_list_ = nullptr;   // Monte-Carlo vectors all have one size
    std::vector<Cached>& list_for(size_t rounded) {
        if (last_list_ && last_size_ == rounded) return *last_list_;
        last_list_ = &free_[rounded]; last_size_ = rounded;    // references into an unordered_map stay valid across rehashes
        return *last_list_;
    }
    static size_t round_size(size_t b);
    void* carve(size_t rounded);
};

// pinned staging buffer for H2D / D2H (replaces the pageable cuMemcpy of RVC:457-481)
struct Staging {
    void* host = nullptr;
    size_t bytes = 0;
    void ensure(size_t need);
    void release();
};

// ---------------------------------------------------------------------------------------------------------
// pending-op graph
// ---------------------------------------------------------------------------------------------------------
enum NodeOp : uint8_t {
    N_LEAF = 0,
    // binary: in[0], in[1] (node index or -1 => scalar imm[k])
    N_ADD, N_SUB, N_MUL, N_DIV, N_MIN, N_MAX,
    // unary on in[0]
    N_SQRT, N_EXP, N_LOG, N_SIN, N_COS, N_ABS, N_INV, N_ISNAN,
    N_POW,       // in[0] ^ imm[1]
    N_CHOOSE,    // in[0] >= 0 ? in[1] : in[2]   (in[1], in[2] may be scalars)
    N_CONST      // every element = imm[0]
};
enum NodeState : uint8_t { NS_FREE = 0, NS_LAZY = 1, NS_MAT = 2 };

struct alignas(64) Node {        // exactly one cache line: the cone walk of every flush touches each pending node
    uint8_t op = N_LEAF, state = NS_FREE;
    int32_t in[3] = {-1, -1, -1};
    float imm[3] = {0.f, 0.f, 0.f};
    uint32_t ext_refs = 0;      // handles held by the caller
    uint32_t int_refs = 0;      // operand slots of pending (lazy) nodes that reference this node
    uint32_t gen = 1;
    // scratch used during one flush
    uint32_t epoch = 0;
    int32_t local = -1;
    int64_t n = 0;
    float* buf = nullptr;       // device vector when NS_MAT
};
static_assert(sizeof(Node) == 64, "Node is one cache line");

struct Options {
    int64_t flush_threshold = 4096;
    bool fuse = true;
    bool profile = false;       // time every interpreter launch with CUDA events (benchmarks)
    // interpreter scheduling knobs (see codegen.cpp: Gen::schedule / Gen::launch)
    int ring_max = TAPE_MAX_RING;   // TMA ring slots per warp, upper bound
    int ring_min = 2;               // ... lower bound when shared memory is tight
    int target_ctas = 0;            // > 0: CTAs per SM the slot budget aims for (tests / tuning); 0: as many as the vector needs, up to the register limit
    int horizon = 96;               // uses of a leaf further apart than this many instructions are separate TMA copies
    bool pipeline = true;           // cross-chunk prefetch (prologue + T_LOADN)
    int max_sets = 1;               // slot sets per warp (cross-chunk prefetch depth), upper bound; measured: >1 costs occupancy and does not pay
    bool zero_copy_reduce = true;   // reductions publish their result through mapped pinned memory (single-rank runs)
    bool leaf_reduce_kernel = true; // reductions of a materialised vector use the streaming kernel, not the interpreter
    bool p2p_reduce = true;         // sharded runs, exchange == 1: partials exchanged inside the kernel over NVLink peer memory
    int exchange = 2;               // how the ranks of a sharded run exchange reduction partials: 2 a table in shared host memory that every
                                    // rank's kernel writes and every rank's host merges (no kernel waits for a peer); 1 inside the kernel over
                                    // NVLink peer memory (a rendez-vous of the kernels); 0 ncclAllGather behind the kernel
    double exchange_timeout_s = 120.0;   // a peer that has not delivered its partial after this long is taken for dead (FMC_ERR_COMM)
    int cta_warps = 4;              // warps per interpreter CTA (1..TAPE_MAX_WARPS)
    int tape_elems = 0;             // chunk geometry: elements per lane, 16 / 8 / 4; 0 = chosen per launch from the vector length
    int min_warps = 5;              // ... the largest geometry that still gives every SM this many warps of work (measured on the LMM step: 16-element
                                    // chunks win down to ~400 k paths, 8-element chunks below)
    bool fuse_ops = true;           // peephole fusion of the abstract code (MULADD_II, ACCUM_S, ADDPROD)
    bool fuse_ops2 = true;          // ... and the one-dispatch forms on top of it (RATIOACC, AXPYST, ADDAFFDISC with its slot's reload)
    int brownian_blocks_per_sm = 0; // blocks per SM the Brownian generator sizes its grid for (0: what the occupancy query reports, one wave)
    int grid_limit = 0;             // > 0: cap the interpreter grid (tests: many chunks per warp at small sizes)
    int max_regs = 8;               // register-file slots the code generator may use, <= TAPE_REGS. Measured on the LMM step: 16 slots
                                    // let a few kernels drop to one CTA per SM (8.36 ms simulation); 8: 8.14 ms, 4: 8.00 ms but more spills
    bool tape_upload_stream = true; // long tapes reach the device through the copy stream, ahead of the kernels queued on the compute stream
    bool tape_cache = true;         // replay the launches of a cone whose structure was lowered before (codegen.cpp)
    bool regression_float_products = true;    // normal equations from the float products fl32(a*b) like RandomVariableFromFloatArray (the reference's
                                    // float class: coefficients within 1e-5 of it even for ill-conditioned bases; compute-bound, 45-49 % of the
                                    // roofline at k = 6..8). false: from the exact products (RandomVariableFromDoubleArray's value for the same
                                    // inputs; sums differ by ~2^-24 / sqrt(n) relative; HBM-bound, 80 % of the roofline)
    int window_levels = 3;          // > 0: a flush without a reduction is cut into windows of this many dependency levels of still-referenced
                                    // values (time steps of a simulation) and each window is emitted chain by chain (component-major) with the
                                    // chains' running state kept on chip, so a value a step stores is not read back from HBM by the next step
                                    // inside the window (codegen.cpp: Runtime::run_cone)
    int window_elems = 8;           // chunk geometry of window kernels (0: the rule of tape_elems / min_warps)
    int window_cta_warps = 16;      // warps per CTA of window kernels (0: cta_warps)
    int window_reduce_min = 2048;   // a reduction whose target is pending first sends the still-referenced pending values below it through the
                                    // windows when more than this many nodes are pending (the tail of a simulation that the first valuation reads)
    int window_ring_extra = 3;      // ring slots of a window kernel beyond the ones its long-lived leaves occupy
    bool batch_reduce = true;       // getAverage() of a vector that one flush materialised together with others: the sums of all of them in one
                                    // launch, the others' results kept for the calls that follow (Runtime::reduce_batch)
};

struct Stats {
    uint64_t n_ops = 0, n_kernels = 0, n_tape_kernels = 0, n_tape_instr = 0, n_stored = 0, n_fused = 0, n_flushes = 0;
    uint64_t h2d = 0, d2h = 0;
};

// host-side time spent per phase (microseconds since the last reset); read with fmc_get_option("host_us_*")
struct HostProfile { double codegen = 0, launch = 0, sync = 0, upload = 0, upload_wait = 0; };   // upload: whole host->device calls; upload_wait: of that, waiting for a free staging chunk

struct Operand { int32_t node; float imm; };   // node < 0 => scalar

struct ReduceSpec {
    int mode = RM_NONE;          // tape reduce mode
    double param = 0.0;
    int32_t weight = -1;         // node index of the weight vector (RM_DOT / RM_WSQ)
};

// The runtime lock. Critical sections are ~100 ns (one recorded op), so waiters spin: test-and-test-and-set (a waiter only
// reads the flag until it sees it free, so the holder's unlock does not fight for the cache line), yielding the core when the
// holder is in a long operation (an upload, a Brownian generation).
class RuntimeMutex {
    std::atomic<int> state_{0};
public:
    bool try_lock() { return state_.load(std::memory_order_relaxed) == 0 && state_.exchange(1, std::memory_order_acquire) == 0; }
    void lock() {
        for (unsigned n = 0; !try_lock(); ) {
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
            if (++n > (1u << 14)) { std::this_thread::yield(); n = 1u << 13; }
        }
    }
    void unlock() { state_.store(0, std::memory_order_release); }
};
using RuntimeLock = std::unique_lock<RuntimeMutex>;

class Runtime {
public:
    static Runtime& get();
    RuntimeMutex mu;

    // lifecycle
    bool initialized = false;
    int device = -1, sm_count = 0;
    size_t smem_per_sm = 0, smem_per_cta_max = 0;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    // host -> device uploads run on their own stream so that they overlap the kernels already queued (the compute stream
    // waits for the copy's event before it goes on); STAGING_SLOTS pinned chunks let the host run ahead of the copies
    static constexpr int STAGING_SLOTS = 8;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_copy[STAGING_SLOTS] = {nullptr}, ev_order = nullptr;
    uint64_t settled_stamp = 0;         // pool blocks freed up to this stamp are no longer touched by any queued kernel
    void sync_stream();                 // cudaStreamSynchronize(stream) + settled_stamp update
    void init(int device_index);
    void shutdown();
    void require_init() const;

    // memory
    DevicePool pool;
    Staging staging;
    bool staging_busy[STAGING_SLOTS] = {false};   // a host->device copy out of that staging chunk may still be in flight
    int staging_next = 0;
    void staging_quiesce();                  // wait for those copies (before the staging buffer is reused for something else)
    double* d_partials = nullptr;       // reduction scratch
    unsigned int* d_counter = nullptr;
    double* d_result = nullptr;         // [1024] doubles: [0,4) reduction, [8,40) rank gather, [64,160) regression, [200] path count,
                                        //   [256,356) histogram points, [512,768) order-statistics scratch
    double* h_result = nullptr;         // pinned mirror
    // mapped pinned [TICKET_SLOTS][4]: {count, value, M2, ticket} written by the reduction's last block. Every reduction in
    // flight owns one slot, so several host threads can each wait for their own result (the runtime lock is released while
    // waiting, see Runtime::reduce); with no slot free the reduction falls back to a copy + stream synchronisation.
    static constexpr int TICKET_SLOTS = 64;
    double* h_ticket = nullptr;
    double* h_ticket_dev = nullptr;     // device-side address of h_ticket
    uint64_t ticket_slots_busy = 0;     // bit s: slot s belongs to a waiting reduction
    int reduce_slot = -1;               // slot of the reduction being launched (-1: none)
    double reduce_ticket = 0.0;
    RuntimeLock* held = nullptr;   // the C-ABI call's lock on `mu` (capi.cpp: guarded)
    double last_tape_ticket = 0.0;      // ticket of the last fused chain -> reduce launch (0: none published)
    // Batched averages. A caller that first builds many result vectors and then asks for their averages one by one (finmath-lib's
    // calibration objective: all product values, then value.getAverage() in a second loop) would pay one launch and one blocking
    // round trip per vector. flush_all() remembers which still-referenced vectors it materialised together; the first
    // getAverage() of one of them sums ALL of them in one launch (reduce_kernel.cu: batch_sum_kernel) and keeps the other sums
    // for the calls that follow. Bit-identical to the single-vector kernel; single-rank runs only.
    struct FlushBatch { uint32_t id; std::vector<std::pair<int32_t, uint32_t>> members; };   // {node, generation}
    std::deque<FlushBatch> flush_batches;                                // the most recent ones
    uint32_t flush_batch_seq = 0;
    struct Prefetched { uint32_t gen; double sum; };
    std::unordered_map<int32_t, Prefetched> prefetched;                  // node -> its sum, until the node is freed
    bool reduce_streak = false;         // the last reduction was served by a batch and nothing was recorded since
    double* d_batch_partials = nullptr; unsigned int* d_batch_counters = nullptr;
    double* h_batch = nullptr; double* h_batch_dev = nullptr;            // mapped pinned [BATCH_MAX + 1]
    double batch_ticket = 0.0;
    static uint32_t batch_of(const Node& nd) { uint32_t b; std::memcpy(&b, &nd.imm[2], 4); return b; }      // valid for NS_MAT nodes (imm unused)
    static void set_batch_of(Node& nd, uint32_t b) { std::memcpy(&nd.imm[2], &b, 4); }
    bool reduce_batch(int32_t idx, double out[3]);                       // true: out = {count, sum, 0} of node idx
    int max_grid = 0;

    // graph
    std::vector<Node> nodes;
    std::vector<int32_t> free_nodes;
    std::vector<int32_t> pending;       // lazy nodes in creation order (may contain stale entries)
    uint32_t epoch = 0;
    int64_t n_lazy = 0, n_live_handles = 0;
    Options opt;
    Stats stats;
    HostProfile hostprof;

    // profile mode: event pairs around interpreter launches + their algorithmic bytes
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    size_t prof_used = 0;
    std::vector<uint64_t> prof_launch_bytes;               // algorithmic bytes of each launch since the last read
    uint64_t prof_bytes = 0, prof_launches = 0, prof_touched = 0;   // touched: every vector a kernel reads or writes, re-reads of earlier results included
    void profile_begin();                                  // records the start event of the next launch
    void profile_end(uint64_t algorithmic_bytes, uint64_t touched_bytes = 0);
    void profile_read(double* ms, uint64_t* bytes, uint64_t* launches);   // synchronises, sums, resets

    // handles
    fmc_vec handle_of(int32_t idx) const { return ((uint64_t)nodes[idx].gen << 32) | (uint32_t)(idx + 1); }
    int32_t resolve(fmc_vec h) const;   // throws on invalid

    int32_t new_node();
    int32_t new_leaf(int64_t n);                          // allocates the device buffer, ext_refs = 1
    int32_t record(NodeOp op, int64_t n, Operand a, Operand b = {-1, 0.f}, Operand c = {-1, 0.f});
    void retain(int32_t idx);
    void release_ext(int32_t idx);
    void release_int(int32_t idx);
    void maybe_free(int32_t idx);

    // execution
    void flush_all(bool automatic = false);               // materialise every referenced pending node (automatic: may hold back an incomplete window)
    void materialize(int32_t idx);                        // make node idx NS_MAT
    void run_cone(const std::vector<int32_t>& targets, const ReduceSpec* red);   // core scheduler
    bool run_windows(const std::vector<int32_t>& targets, const std::vector<int32_t>& recorded, bool hold_last);       // a flush in windows of opt.window_levels levels (codegen.cpp)
    int window_levels_now = 0;                            // levels per window in use (<= opt.window_levels; lowered when windows spill, codegen.cpp: run_windows)
    bool windowing = false;                               // run_cone is working through the windows of a flush
    std::unordered_set<const float*> window_stored;       // ... buffers the earlier windows of this flush wrote (traffic accounting)
    int64_t flush_floor = 0;                              // pending nodes the last automatic flush held back: the next one waits for flush_threshold more
    // {count, value, M2} of the LOCAL slice; with several ranks the partials of ALL ranks are left in h_result[4 * r + 0..2]
    // (one ncclAllGather behind the reduction kernel, one copy, one synchronisation)
    // returns true when `out` already is the merged result of all ranks (in-kernel exchange)
    bool reduce(int32_t idx, const ReduceSpec& spec, double out[3]);
    void auto_flush();

    // host copies
    int32_t upload_f64(const double* h, int64_t n);
    int32_t upload_f32(const float* h, int64_t n);
    // asynchronous upload from pinned host memory (fmc_host_alloc): DMA of the doubles on the copy stream into a ring of device
    // chunks, (float) cast on the device. The host buffer is read AFTER the call returns.
    int32_t upload_f64_pinned(const double* h, int64_t n);
    void* host_alloc(size_t bytes);
    void host_free(void* p);
    bool is_pinned(const void* p, size_t bytes) const;
    std::map<const char*, size_t> pinned_;               // base -> size of the fmc_host_alloc allocations
    double* d_upload = nullptr;                           // STAGING_SLOTS chunks of doubles on the device
    void download_f32(int32_t idx, float* h, int64_t n);
    void download_f64(int32_t idx, double* h, int64_t n);

    // comm (NCCL, loaded with dlopen; see comm.cpp)
    int comm_rank = 0, comm_size = 1;
    void* nccl_comm = nullptr;
    // peer-memory exchange of reduction partials (reduce_common.cuh): own table + the peers' tables mapped through cudaIpc
    double* xtable = nullptr;
    double* peer_tables[XMAX_RANKS] = {nullptr};
    bool p2p_ready = false;
    // exchange through shared host memory (POSIX shm, registered with CUDA): [XSLOTS][XMAX_RANKS][4] doubles
    double* xhost = nullptr;            // this process's mapping
    double* xhost_dev = nullptr;        // device-side address of the same memory
    size_t xhost_bytes = 0;
    bool xhost_ready = false;
    bool last_tape_xhost = false;       // the last fused chain -> reduce launch published into the host table
    bool use_xhost() const { return comm_size > 1 && xhost_ready && opt.exchange == 2; }
    bool use_p2p() const { return comm_size > 1 && p2p_ready && opt.exchange == 1 && opt.p2p_reduce; }
    double xticket = 0.0;                                 // same sequence on every rank: reset by comm_init
    void fill_exchange(Exchange& x, double* ticket);      // parameters of the next reduction kernel
    void allreduce_sum(double* dev, int count);           // in place on the compute stream
    void allgather(const double* dev_send, double* dev_recv, int count_per_rank);
    void allreduce_minmax(double* dev, int count, bool is_max);
};

// codegen.cpp
void tape_cache_clear();
void tape_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* entries);

// comm.cpp
void comm_get_unique_id(char* id);
void comm_init(Runtime& rt, int rank, int nranks, const char* id);
void comm_destroy(Runtime& rt);

// brownian.cpp
void brownian_generate(Runtime& rt, int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1,
                       const double* sqrt_dt, int32_t* out_nodes);
void mt19937_raw(Runtime& rt, int seed_mode, int64_t seed, uint64_t skip, int64_t count, uint32_t* host_out);
void brownian_release_caches(Runtime& rt);

}  // namespace fmc
