// runtime.cpp — device lifecycle, handles, recording of operations, host<->device copies.
// (the scheduler / code generator that turns pending nodes into interpreter tapes is in codegen.cpp)
#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <thread>
#if defined(__SSE2__)
#include <immintrin.h>
#endif

#include "runtime.h"

namespace fmc {

Runtime& Runtime::get() {
    static Runtime* rt = new Runtime();   // intentionally leaked: no static-destruction order problems at exit
    return *rt;
}

void Runtime::require_init() const {
    if (!initialized) fail(FMC_ERR_NOT_INIT, "fmcuda runtime not initialised: call fmc_init() (a CUDA device is required; there is no CPU fallback)");
}

void Runtime::init(int device_index) {
    if (initialized) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        fail(FMC_ERR_NOT_INIT, "no CUDA device available (%s); this backend has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    // negative index counts from the end, -1 = last device (RandomVariableCuda.java:161,177)
    int dev = device_index >= 0 ? device_index : count + device_index;
    if (dev < 0 || dev >= count) fail(FMC_ERR_INVALID, "device index %d out of range (device count %d)", device_index, count);
    FMC_CUDA(cudaSetDevice(dev));
    FMC_CUDA(cudaGetDeviceProperties(&prop, dev));
    device = dev;
    sm_count = prop.multiProcessorCount;
    smem_per_sm = prop.sharedMemPerMultiprocessor;
    FMC_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    FMC_CUDA(cudaEventCreate(&ev_start));
    FMC_CUDA(cudaEventCreate(&ev_stop));
    FMC_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < STAGING_SLOTS; i++) { FMC_CUDA(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming)); staging_busy[i] = false; }
    FMC_CUDA(cudaEventCreateWithFlags(&ev_order, cudaEventDisableTiming));
    settled_stamp = 0;
    FMC_CUDA(tape_kernel_setup(&smem_per_cta_max));
    max_grid = sm_count * 16;
    FMC_CUDA(cudaMalloc(&d_partials, sizeof(double) * 128 * (size_t)max_grid));
    FMC_CUDA(cudaMalloc(&d_counter, sizeof(unsigned int) * 4));
    FMC_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int) * 4, stream));     // on the compute stream: ordered before the first kernel that counts
    FMC_CUDA(cudaMalloc(&d_result, sizeof(double) * 1024));
    FMC_CUDA(cudaMallocHost(&h_result, sizeof(double) * 1024));
    FMC_CUDA(cudaHostAlloc(&h_ticket, sizeof(double) * 4 * TICKET_SLOTS, cudaHostAllocMapped));
    FMC_CUDA(cudaHostGetDevicePointer(&h_ticket_dev, h_ticket, 0));
    for (int i = 0; i < 4 * TICKET_SLOTS; i++) h_ticket[i] = 0.0;
    reduce_ticket = 0.0; ticket_slots_busy = 0; reduce_slot = -1;
    nodes.reserve(1 << 16);
    initialized = true;
}

void Runtime::shutdown() {
    if (!initialized) return;
    cudaStreamSynchronize(stream);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    for (int i = 0; i < STAGING_SLOTS; i++) staging_busy[i] = false;
    comm_destroy(*this);
    brownian_release_caches(*this);
    for (auto& nd : nodes) {
        if (nd.state == NS_MAT && nd.buf) pool.free(nd.buf);
        nd = Node{};
    }
    nodes.clear(); free_nodes.clear(); pending.clear();
    n_lazy = 0; n_live_handles = 0;
    pool.purge();
    staging.release();
    tape_kernel_teardown();
    tape_cache_clear();
    if (d_upload) { cudaFree(d_upload); d_upload = nullptr; }
    if (d_batch_partials) { cudaFree(d_batch_partials); d_batch_partials = nullptr; }
    if (d_batch_counters) { cudaFree(d_batch_counters); d_batch_counters = nullptr; }
    if (h_batch) { cudaFreeHost(h_batch); h_batch = nullptr; h_batch_dev = nullptr; }
    flush_batches.clear(); prefetched.clear(); reduce_streak = false; batch_ticket = 0.0;
    for (auto& kv : pinned_) cudaFreeHost((void*)kv.first);
    pinned_.clear();
    cudaFree(d_partials); cudaFree(d_counter); cudaFree(d_result); cudaFreeHost(h_result); cudaFreeHost(h_ticket);
    d_partials = nullptr; d_counter = nullptr; d_result = nullptr; h_result = nullptr; h_ticket = nullptr; h_ticket_dev = nullptr;
    for (auto& pe : prof_events) { cudaEventDestroy(pe.first); cudaEventDestroy(pe.second); }
    prof_events.clear(); prof_used = 0;
    cudaEventDestroy(ev_start); cudaEventDestroy(ev_stop); for (int i = 0; i < STAGING_SLOTS; i++) { cudaEventDestroy(ev_copy[i]); ev_copy[i] = nullptr; }
    cudaEventDestroy(ev_order); ev_order = nullptr;
    if (copy_stream) { cudaStreamDestroy(copy_stream); copy_stream = nullptr; }
    cudaStreamDestroy(stream);
    stream = nullptr;
    initialized = false;
}

// ---------------------------------------------------------------------------------------------------------
// handles and nodes
// ---------------------------------------------------------------------------------------------------------
int32_t Runtime::resolve(fmc_vec h) const {
    const uint32_t lo = (uint32_t)(h & 0xffffffffu), gen = (uint32_t)(h >> 32);
    if (lo == 0 || lo > nodes.size()) fail(FMC_ERR_INVALID, "invalid vector handle 0x%llx", (unsigned long long)h);
    const int32_t idx = (int32_t)lo - 1;
    const Node& nd = nodes[idx];
    if (nd.state == NS_FREE || nd.gen != gen || nd.ext_refs == 0)
        fail(FMC_ERR_INVALID, "stale or released vector handle 0x%llx", (unsigned long long)h);
    return idx;
}

int32_t Runtime::new_node() {
    int32_t idx;
    if (!free_nodes.empty()) { idx = free_nodes.back(); free_nodes.pop_back(); }
    else { nodes.emplace_back(); idx = (int32_t)nodes.size() - 1; }
    Node& nd = nodes[idx];
    const uint32_t gen = nd.gen;
    nd = Node{};
    nd.gen = gen;
    return idx;
}

int32_t Runtime::new_leaf(int64_t n) {
    if (n < 0) fail(FMC_ERR_INVALID, "negative vector size %lld", (long long)n);
    const int32_t idx = new_node();
    Node& nd = nodes[idx];
    nd.op = N_LEAF; nd.state = NS_MAT; nd.n = n; nd.ext_refs = 1;
    nd.buf = (float*)pool.alloc(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
    n_live_handles++;
    return idx;
}

int32_t Runtime::record(NodeOp op, int64_t n, Operand a, Operand b, Operand c) {
    const int32_t idx = new_node();
    Node& nd = nodes[idx];
    nd.op = op; nd.state = NS_LAZY; nd.n = n; nd.ext_refs = 1;
    const Operand ops[3] = {a, b, c};
    for (int k = 0; k < 3; k++) {
        nd.in[k] = ops[k].node;
        nd.imm[k] = ops[k].imm;
        if (ops[k].node >= 0) nodes[ops[k].node].int_refs++;
    }
    n_lazy++; n_live_handles++;
    pending.push_back(idx);
    stats.n_ops++;
    reduce_streak = false;
    return idx;
}

void Runtime::retain(int32_t idx) { nodes[idx].ext_refs++; }

void Runtime::maybe_free(int32_t first) {
    // iterative cascade: freeing a lazy node releases its operands (explicit stack: chains of a few thousand pending nodes are
    // normal; the common cases — nothing to free, or one node — never touch the heap)
    {
        const Node& nd = nodes[first];
        if (nd.state == NS_FREE || nd.ext_refs != 0 || nd.int_refs != 0) return;
    }
    int32_t small[16];
    int n_small = 0;
    std::vector<int32_t> big;
    auto push = [&](int32_t v) { if (n_small < 16) small[n_small++] = v; else big.push_back(v); };
    auto pop = [&]() -> int32_t { if (!big.empty()) { const int32_t v = big.back(); big.pop_back(); return v; } return small[--n_small]; };
    push(first);
    while (n_small > 0 || !big.empty()) {
        const int32_t idx = pop();
        Node& nd = nodes[idx];
        if (nd.state == NS_FREE || nd.ext_refs != 0 || nd.int_refs != 0) continue;
        if (nd.state == NS_LAZY) {
            n_lazy--;
            for (int k = 0; k < 3; k++) if (nd.in[k] >= 0) {
                Node& in = nodes[nd.in[k]];
                if (in.int_refs > 0) in.int_refs--;
                if (in.int_refs == 0 && in.ext_refs == 0) push(nd.in[k]);
            }
        } else if (nd.buf) {
            pool.free(nd.buf);
            if (!prefetched.empty()) prefetched.erase(idx);
        }
        nd.buf = nullptr; nd.state = NS_FREE; nd.gen++;
        if (nd.gen == 0) nd.gen = 1;
        free_nodes.push_back(idx);
    }
}

void Runtime::release_ext(int32_t idx) {
    Node& nd = nodes[idx];
    if (nd.ext_refs == 0) fail(FMC_ERR_INVALID, "release of a handle with zero references");
    nd.ext_refs--;
    if (nd.ext_refs == 0) { n_live_handles--; maybe_free(idx); }
}

void Runtime::release_int(int32_t idx) {
    Node& nd = nodes[idx];
    if (nd.int_refs > 0) nd.int_refs--;
    maybe_free(idx);
}

void Runtime::auto_flush() {
    if (!opt.fuse || n_lazy > opt.flush_threshold + flush_floor) flush_all(true);
}

void Runtime::flush_all(bool automatic) {
    if (n_lazy == 0) { pending.clear(); flush_floor = 0; return; }
    std::vector<int32_t> targets;
    targets.reserve(pending.size());
    for (int32_t idx : pending) {
        const Node& nd = nodes[idx];
        if (nd.state == NS_LAZY && nd.ext_refs > 0) targets.push_back(idx);
    }
    const bool windows = opt.fuse && opt.window_levels > 0 && targets.size() > 1;
    // an automatic flush leaves its last, incomplete window pending (run_windows) unless the graph has grown far beyond the threshold
    const bool hold = windows && automatic && n_lazy <= 4 * opt.flush_threshold;
    static thread_local std::vector<int32_t> order;      // the pending list of this flush (run_windows walks it; run_cone may record nothing, but stay safe)
    order.clear();
    if (!hold) { order.swap(pending); flush_floor = 0; }
    if (targets.empty()) return;
    if (windows) {
        const bool held = run_windows(targets, hold ? pending : order, hold);
        if (hold) {
            if (held) {
                size_t k = 0;
                for (int32_t idx : pending) if (nodes[idx].state == NS_LAZY) pending[k++] = idx;
                pending.resize(k);
                flush_floor = n_lazy;
            } else { pending.clear(); flush_floor = 0; }
        }
    } else run_cone(targets, nullptr);
    // the vectors this flush materialised together (see runtime.h: batched averages)
    if (opt.batch_reduce && comm_size == 1 && targets.size() >= 2) {
        FlushBatch fb;
        fb.id = ++flush_batch_seq;
        if (fb.id == 0) fb.id = ++flush_batch_seq;
        fb.members.reserve(targets.size());
        for (int32_t t : targets) {
            Node& nd = nodes[t];
            if (nd.state != NS_MAT || nd.ext_refs == 0 || batch_of(nd) == fb.id) continue;
            set_batch_of(nd, fb.id);
            fb.members.emplace_back(t, nd.gen);
        }
        if (fb.members.size() >= 2) {
            flush_batches.push_back(std::move(fb));
            if (flush_batches.size() > 64) flush_batches.pop_front();
        }
    }
}

// The sums of the still-referenced, unconsumed vectors that were materialised by the same flush as node idx, in one launch.
// Returns false (nothing done) when idx has no such siblings.
bool Runtime::reduce_batch(int32_t idx, double out[3]) {
    const Node& me = nodes[idx];
    const uint32_t id = batch_of(me);
    if (id == 0) return false;
    FlushBatch* fb = nullptr;
    for (auto it = flush_batches.rbegin(); it != flush_batches.rend(); ++it) if (it->id == id) { fb = &*it; break; }
    if (!fb) return false;
    // candidates in recording order, starting with idx itself (at most BATCH_MAX: the ones recorded right after idx first)
    std::vector<int32_t> cand;
    cand.reserve(std::min<size_t>(fb->members.size(), (size_t)BATCH_MAX));
    size_t at = fb->members.size();
    for (size_t k = 0; k < fb->members.size(); k++) if (fb->members[k].first == idx && fb->members[k].second == me.gen) { at = k; break; }
    if (at == fb->members.size()) return false;
    for (size_t d = 0; d < fb->members.size() && cand.size() < (size_t)BATCH_MAX; d++) {
        const auto& m = fb->members[(at + d) % fb->members.size()];
        const Node& nd = nodes[m.first];
        if (nd.state != NS_MAT || nd.gen != m.second || nd.ext_refs == 0 || nd.int_refs != 0 || nd.n != me.n || !nd.buf) continue;
        if (batch_of(nd) != id) continue;
        auto pf = prefetched.find(m.first);
        if (pf != prefetched.end() && pf->second.gen == nd.gen) continue;
        cand.push_back(m.first);
    }
    if (cand.size() < 2 || cand[0] != idx) return false;
    const int64_t n = me.n;
    const int64_t tiles = (n + reduce_tile_elems() - 1) / reduce_tile_elems();
    int B = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)sm_count * 8));    // the grid Runtime::reduce gives launch_reduce
    B = std::min(B, max_grid);
    if (opt.grid_limit > 0) B = std::min(B, opt.grid_limit);
    if (!d_batch_partials) {
        FMC_CUDA(cudaMalloc(&d_batch_partials, sizeof(double) * 2 * (size_t)BATCH_MAX * (size_t)max_grid));
        FMC_CUDA(cudaMalloc(&d_batch_counters, sizeof(unsigned int) * (BATCH_MAX + 1)));
        FMC_CUDA(cudaMemsetAsync(d_batch_counters, 0, sizeof(unsigned int) * (BATCH_MAX + 1), stream));
        FMC_CUDA(cudaHostAlloc(&h_batch, sizeof(double) * (BATCH_MAX + 1), cudaHostAllocMapped));
        FMC_CUDA(cudaHostGetDevicePointer(&h_batch_dev, h_batch, 0));
        for (int i = 0; i <= BATCH_MAX; i++) h_batch[i] = 0.0;
    }
    static thread_local BatchSumParams P;
    P.n = n; P.k = (int)cand.size(); P.blocks_per_vec = B;
    P.partials = d_batch_partials; P.counters = d_batch_counters; P.host_out = h_batch_dev;
    P.ticket = (batch_ticket += 1.0);
    for (size_t k = 0; k < cand.size(); k++) P.x[k] = nodes[cand[k]].buf;
    if (opt.profile) profile_begin();
    const auto t_launch0 = std::chrono::steady_clock::now();
    FMC_CUDA(launch_batch_sum(P, stream));
    hostprof.launch += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_launch0).count();
    if (opt.profile) profile_end(4ull * (uint64_t)n * (uint64_t)cand.size());
    stats.n_kernels++; stats.n_flushes++;
    const auto t_sync0 = std::chrono::steady_clock::now();
    const uint64_t stamp_at_launch = pool.free_stamp;
    volatile double* h = h_batch;
    unsigned spins = 0;
    auto t_query = t_sync0;
    while (h[BATCH_MAX] != P.ticket) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((++spins & 0x3ffu) == 0u) {
            const auto now = std::chrono::steady_clock::now();
            if (now - t_query < std::chrono::microseconds(200)) continue;
            t_query = now;
            const cudaError_t q = cudaStreamQuery(stream);
            if (q == cudaSuccess) { if (h[BATCH_MAX] == P.ticket) break; fail(FMC_ERR_CUDA, "batched reduction finished without publishing its results"); }
            if (q != cudaErrorNotReady) FMC_CUDA(q);
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    for (size_t k = 1; k < cand.size(); k++) prefetched[cand[k]] = Prefetched{nodes[cand[k]].gen, h[k]};
    out[0] = (double)n; out[1] = h[0]; out[2] = 0.0;
    settled_stamp = std::max(settled_stamp, stamp_at_launch);
    hostprof.sync += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_sync0).count();
    stats.d2h += 8 * (uint64_t)cand.size();
    return true;
}

void Runtime::materialize(int32_t idx) {
    if (nodes[idx].state == NS_MAT) return;
    std::vector<int32_t> t{idx};
    run_cone(t, nullptr);
}

void Runtime::profile_begin() {
    if (prof_used == prof_events.size()) {
        cudaEvent_t a, b;
        FMC_CUDA(cudaEventCreate(&a));
        FMC_CUDA(cudaEventCreate(&b));
        prof_events.emplace_back(a, b);
    }
    FMC_CUDA(cudaEventRecord(prof_events[prof_used].first, stream));
}
void Runtime::profile_end(uint64_t algorithmic_bytes, uint64_t touched_bytes) {
    FMC_CUDA(cudaEventRecord(prof_events[prof_used].second, stream));
    prof_used++;
    prof_launch_bytes.push_back(algorithmic_bytes);
    prof_bytes += algorithmic_bytes;
    prof_touched += touched_bytes ? touched_bytes : algorithmic_bytes;
    prof_launches++;
}
void Runtime::profile_read(double* ms, uint64_t* bytes, uint64_t* launches) {
    FMC_CUDA(cudaStreamSynchronize(stream));
    double total = 0.0;
    static const char* dump = std::getenv("FMC_PROFILE_DUMP");       // development aid: one line per launch (microseconds, algorithmic bytes)
    FILE* df = dump ? std::fopen(dump, "a") : nullptr;
    for (size_t i = 0; i < prof_used; i++) {
        float t = 0.f;
        FMC_CUDA(cudaEventElapsedTime(&t, prof_events[i].first, prof_events[i].second));
        total += t;
        if (df) std::fprintf(df, "%zu %.2f %llu\n", i, 1e3 * (double)t, (unsigned long long)(i < prof_launch_bytes.size() ? prof_launch_bytes[i] : 0));
    }
    if (df) { std::fprintf(df, "# read\n"); std::fclose(df); }
    prof_launch_bytes.clear();
    if (ms) *ms = total;
    if (bytes) *bytes = prof_bytes;
    if (launches) *launches = prof_launches;
    prof_used = 0; prof_bytes = 0; prof_launches = 0; prof_touched = 0;
}

// ---------------------------------------------------------------------------------------------------------
// host <-> device copies through pinned staging (double buffered, chunked)
//
// The (float) cast of createRandomVariable(time, double[]) (RandomVariableCuda.java:768-774) and the widening of
// getRealizations() (RVC:776-782) are host loops over every element; at 1M paths x 80 vectors they, not PCIe, bound
// the end-to-end time when run on one core. A small persistent worker pool converts each staging chunk in parallel
// while the previous chunk's cudaMemcpyAsync is in flight.
// ---------------------------------------------------------------------------------------------------------
namespace {

class HostWorkers {
public:
    static HostWorkers& get() { static HostWorkers* w = new HostWorkers(); return *w; }   // leaked like Runtime
    // fn(begin, end) over [0, n) split into contiguous blocks; returns when all blocks are done
    void parallel_for(int64_t n, const std::function<void(int64_t, int64_t)>& fn) {
        // blocks of 32 Ki elements, four per thread at most, handed out one at a time: a worker that wakes up late (a futex wake-up on a
        // virtual CPU takes tens of microseconds, a third of what a 4 MiB chunk costs) takes fewer blocks instead of holding up the call
        const int64_t min_block = 1 << 15;
        const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_per_thread_ * ((int64_t)threads_.size() + 1), n / min_block));
        if (parts <= 1 || threads_.empty()) { fn(0, n); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; parts_ = parts; next_ = 0; pending_.store(parts, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        work();                                  // the caller takes blocks too
        // the last blocks are a few microseconds from done: look before sleeping
        for (int spin = 0; spin < 20000 && pending_.load(std::memory_order_acquire) != 0; spin++) cpu_relax();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_.load(std::memory_order_acquire) == 0; });
        fn_ = nullptr;
    }
private:
    HostWorkers() {
        unsigned hc = std::thread::hardware_concurrency();
        // measured on the 16-vCPU single-B200 box, end-to-end LMM step with blocks handed out dynamically and spinning workers
        // (gpurun_out/r5e.log): 6 threads 27-28 ms, 8: 25-26, 12: 23.6, 16: 22.7-23.2 (with one static block per thread and
        // sleeping workers 8 threads were the best: 26 ms). One process per GPU shares the host: the cores are divided by the number
        // of local ranks (torchrun's LOCAL_WORLD_SIZE). Measured at 4 ranks on 32 vCPUs: 4 threads 52.0 ms, 8: 44.6 ms, 12: 44.2 ms;
        // at 2 ranks on 24 vCPUs (r5j.log): 8 threads 28.9 ms, 12 threads 28.7 ms, 12 spinning 26.4 ms. Workers spin only where a rank
        // has 12 or more cores to itself (8 ranks on 32 vCPUs would fill every core with spinning threads).
        unsigned local_ranks = 1;
        if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) local_ranks = (unsigned)std::max(1, std::atoi(e));
        const unsigned per_rank = (hc ? hc : 8u) / local_ranks;
        int want = (int)std::min<unsigned>(std::max<unsigned>(per_rank, 2u), 16u) - 1;   // FMC_HOST_THREADS overrides
        if (const char* e = std::getenv("FMC_HOST_THREADS")) want = std::max(0, std::atoi(e) - 1);
        spin_us_ = per_rank >= 12 ? 500 : 0;                                                               // FMC_HOST_SPIN_US overrides
        if (const char* e = std::getenv("FMC_HOST_SPIN_US")) spin_us_ = std::max(0, std::atoi(e));
        if (const char* e = std::getenv("FMC_HOST_BLOCKS_PER_THREAD")) blocks_per_thread_ = std::max(1, std::atoi(e));
        for (int i = 0; i < want; i++) threads_.emplace_back([this] { loop(); });
        for (auto& t : threads_) t.detach();
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // uploads come in bursts (one vector per time step of a simulation being recorded): a worker keeps looking for the next
            // call for spin_us_ before it goes to sleep (0 where the ranks of a host have fewer than 12 cores each: see the constructor).
            if (spin_us_ > 0) {
                const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us_);
                while (gen_.load(std::memory_order_acquire) == seen) {
                    for (int k = 0; k < 64 && gen_.load(std::memory_order_relaxed) == seen; k++) cpu_relax();
                    if (std::chrono::steady_clock::now() >= t_end) break;
                }
            }
            if (gen_.load(std::memory_order_acquire) == seen) {
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1, std::memory_order_acq_rel);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_acq_rel);
            }
            seen = gen_.load(std::memory_order_acquire);
            work();
        }
    }
    void work() {
        for (;;) {
            int part; int64_t n; int parts; const std::function<void(int64_t, int64_t)>* fn;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!fn_ || next_ >= parts_) return;
                part = next_++; n = n_; parts = parts_; fn = fn_;
            }
            const int64_t per = (n + parts - 1) / parts;
            const int64_t b = std::min<int64_t>(n, per * part), e = std::min<int64_t>(n, b + per);
            if (b < e) (*fn)(b, e);
            if (pending_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lk(mu_);
                done_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int64_t, int64_t)>* fn_ = nullptr;
    int64_t n_ = 0;
    int parts_ = 0, next_ = 0;
    std::atomic<int> pending_{0}, sleepers_{0};
    std::atomic<uint64_t> gen_{0};
    int spin_us_ = 0, blocks_per_thread_ = 4;
};

constexpr size_t kChunkElems = 4u << 20;   // 16 MiB of floats per staging half

// (float)d[i] into the pinned staging buffer. The staging half is written once and then read by the DMA engine only, so
// on x86-64 the stores are streaming (non-temporal): no read-for-ownership of the destination lines, which is a quarter
// of this loop's memory traffic. cvtpd2ps rounds to nearest even like Java's (float) cast.
#if defined(__x86_64__) && defined(__GNUC__)
// the same loop with 256-bit loads (chosen at run time when the CPU has AVX: half the instructions per element)
__attribute__((target("avx"))) static void cast_to_staging_avx(float* dst, const double* src, int64_t n) {
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31u)) { dst[i] = (float)src[i]; i++; }
    for (; i + 8 <= n; i += 8) {
        const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i)), hi = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
        _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));
    }
    for (; i < n; i++) dst[i] = (float)src[i];
    _mm_sfence();
}
static const bool host_has_avx = [] {
    if (const char* e = std::getenv("FMC_HOST_AVX")) return std::atoi(e) != 0 && __builtin_cpu_supports("avx");
    return (bool)__builtin_cpu_supports("avx");
}();
#define FMC_HAVE_AVX_CAST 1
#endif

inline void cast_to_staging(float* dst, const double* src, int64_t n) {
#if defined(FMC_HAVE_AVX_CAST)
    if (host_has_avx) { cast_to_staging_avx(dst, src, n); return; }
#endif
#if defined(__SSE2__)
    int64_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15u)) { dst[i] = (float)src[i]; i++; }
    for (; i + 4 <= n; i += 4) {
        const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(src + i)), hi = _mm_cvtpd_ps(_mm_loadu_pd(src + i + 2));
        _mm_stream_ps(dst + i, _mm_movelh_ps(lo, hi));
    }
    for (; i < n; i++) dst[i] = (float)src[i];
    _mm_sfence();
#else
    for (int64_t i = 0; i < n; i++) dst[i] = (float)src[i];
#endif
}
inline void cast_to_staging(float* dst, const float* src, int64_t n) { std::memcpy(dst, src, sizeof(float) * (size_t)n); }

}  // namespace

constexpr size_t kUploadChunk = 1u << 20;  // 4 MiB of floats per staging chunk on the upload path (STAGING_SLOTS of them)

template <typename T>
static int32_t upload_impl(Runtime& rt, const T* h, int64_t n) {
    if (n > 0 && !h) fail(FMC_ERR_INVALID, "null host pointer");
    if (n < 0) fail(FMC_ERR_INVALID, "negative vector size %lld", (long long)n);
    struct Timer {
        double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        explicit Timer(double& a) : acc(a) {}
        ~Timer() { acc += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
    } timer(rt.hostprof.upload);
    // The copy runs on the copy stream, concurrently with the kernels already queued on the compute stream. Its destination
    // must therefore be memory no queued kernel still touches: a block freed before the last point the host saw the compute
    // stream idle or a reduction finish, or fresh memory.
    { const uint64_t stamp = rt.pool.free_stamp; if (cudaStreamQuery(rt.stream) == cudaSuccess) rt.settled_stamp = std::max(rt.settled_stamp, stamp); else cudaGetLastError(); }
    bool settled = true;
    float* dst = (float*)rt.pool.alloc_settled(sizeof(float) * (size_t)std::max<int64_t>(n, 1), rt.settled_stamp, &settled);
    int32_t idx;
    try {
        idx = rt.new_node();
    } catch (...) { rt.pool.free(dst); throw; }
    {
        Node& nd = rt.nodes[idx];
        nd.op = N_LEAF; nd.state = NS_MAT; nd.n = n; nd.ext_refs = 1; nd.buf = dst;
        rt.n_live_handles++;
    }
    static_assert(2 * kChunkElems >= Runtime::STAGING_SLOTS * kUploadChunk, "the staging buffer holds all upload chunks");
    rt.staging.ensure(2 * kChunkElems * sizeof(float));
    int64_t off = 0;
    try {
        if (!settled) {                            // recycled too recently: the copy must wait for what is queued
            FMC_CUDA(cudaEventRecord(rt.ev_order, rt.stream));
            FMC_CUDA(cudaStreamWaitEvent(rt.copy_stream, rt.ev_order, 0));
        }
        int last = -1;
        while (off < n) {
            const int64_t m = std::min<int64_t>(kUploadChunk, n - off);
            const int which = rt.staging_next; rt.staging_next = (rt.staging_next + 1) % Runtime::STAGING_SLOTS;
            if (rt.staging_busy[which]) { Timer tw(rt.hostprof.upload_wait); FMC_CUDA(cudaEventSynchronize(rt.ev_copy[which])); rt.staging_busy[which] = false; }
            float* s = (float*)rt.staging.host + (size_t)which * kUploadChunk;
            const T* src = h + off;
            HostWorkers::get().parallel_for(m, [&](int64_t b, int64_t e) { cast_to_staging(s + b, src + b, e - b); });   // RVC:768-774
            FMC_CUDA(cudaMemcpyAsync(dst + off, s, sizeof(float) * (size_t)m, cudaMemcpyHostToDevice, rt.copy_stream));
            FMC_CUDA(cudaEventRecord(rt.ev_copy[which], rt.copy_stream));
            rt.staging_busy[which] = true;
            last = which;
            off += m;
        }
        // no host wait: a staging chunk stays marked busy until its copy's event is seen by its next user. Whatever the
        // compute stream is given from here on runs after the copy.
        if (last >= 0) FMC_CUDA(cudaStreamWaitEvent(rt.stream, rt.ev_copy[last], 0));
    } catch (...) {
        rt.release_ext(idx);
        throw;
    }
    rt.stats.h2d += sizeof(float) * (uint64_t)n;
    return idx;
}
int32_t Runtime::upload_f64(const double* h, int64_t n) { return upload_impl(*this, h, n); }

void* Runtime::host_alloc(size_t bytes) {
    void* p = nullptr;
    FMC_CUDA(cudaMallocHost(&p, std::max<size_t>(bytes, 1)));
    pinned_[(const char*)p] = bytes;
    return p;
}
void Runtime::host_free(void* p) {
    auto it = pinned_.find((const char*)p);
    if (it == pinned_.end()) fail(FMC_ERR_INVALID, "fmc_host_free: not an fmc_host_alloc pointer");
    FMC_CUDA(cudaStreamSynchronize(copy_stream));        // a DMA out of it may still be in flight
    pinned_.erase(it);
    FMC_CUDA(cudaFreeHost(p));
}
bool Runtime::is_pinned(const void* p, size_t bytes) const {
    auto it = pinned_.upper_bound((const char*)p);
    if (it == pinned_.begin()) return false;
    --it;
    return (const char*)p >= it->first && (const char*)p + bytes <= it->first + it->second;
}

int32_t Runtime::upload_f64_pinned(const double* h, int64_t n) {
    if (n < 0) fail(FMC_ERR_INVALID, "negative vector size %lld", (long long)n);
    if (n > 0 && !is_pinned(h, sizeof(double) * (size_t)n)) fail(FMC_ERR_INVALID, "fmc_vec_from_f64_pinned: the buffer does not come from fmc_host_alloc");
    struct Timer {
        double& acc; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        explicit Timer(double& a) : acc(a) {}
        ~Timer() { acc += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
    } timer(hostprof.upload);
    if (!d_upload) FMC_CUDA(cudaMalloc(&d_upload, sizeof(double) * kUploadChunk * STAGING_SLOTS));
    { const uint64_t stamp = pool.free_stamp; if (cudaStreamQuery(stream) == cudaSuccess) settled_stamp = std::max(settled_stamp, stamp); else cudaGetLastError(); }
    bool settled = true;
    float* dst = (float*)pool.alloc_settled(sizeof(float) * (size_t)std::max<int64_t>(n, 1), settled_stamp, &settled);
    int32_t idx;
    try { idx = new_node(); } catch (...) { pool.free(dst); throw; }
    {
        Node& nd = nodes[idx];
        nd.op = N_LEAF; nd.state = NS_MAT; nd.n = n; nd.ext_refs = 1; nd.buf = dst;
        n_live_handles++;
    }
    try {
        if (!settled) {
            FMC_CUDA(cudaEventRecord(ev_order, stream));
            FMC_CUDA(cudaStreamWaitEvent(copy_stream, ev_order, 0));
        }
        int last = -1;
        for (int64_t off = 0; off < n; off += (int64_t)kUploadChunk) {
            const int64_t m = std::min<int64_t>(kUploadChunk, n - off);
            const int which = staging_next; staging_next = (staging_next + 1) % STAGING_SLOTS;
            if (staging_busy[which]) { Timer tw(hostprof.upload_wait); FMC_CUDA(cudaEventSynchronize(ev_copy[which])); staging_busy[which] = false; }
            double* chunk = d_upload + (size_t)which * kUploadChunk;
            FMC_CUDA(cudaMemcpyAsync(chunk, h + off, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, copy_stream));
            FMC_CUDA(launch_cast_f64_f32(chunk, dst + off, m, sm_count, copy_stream));
            FMC_CUDA(cudaEventRecord(ev_copy[which], copy_stream));
            staging_busy[which] = true;
            last = which;
        }
        if (last >= 0) FMC_CUDA(cudaStreamWaitEvent(stream, ev_copy[last], 0));
    } catch (...) {
        release_ext(idx);
        throw;
    }
    stats.h2d += sizeof(double) * (uint64_t)n;
    stats.n_kernels++;
    return idx;
}
int32_t Runtime::upload_f32(const float* h, int64_t n) { return upload_impl(*this, h, n); }

void Runtime::staging_quiesce() {
    for (int w = 0; w < STAGING_SLOTS; w++) if (staging_busy[w]) { FMC_CUDA(cudaEventSynchronize(ev_copy[w])); staging_busy[w] = false; }
}

void Runtime::sync_stream() {
    const uint64_t stamp = pool.free_stamp;          // everything freed so far was last used by work queued before this point
    FMC_CUDA(cudaStreamSynchronize(stream));
    settled_stamp = std::max(settled_stamp, stamp);
}

template <typename T>
static void download_impl(Runtime& rt, int32_t idx, T* h, int64_t n) {
    if (n != rt.nodes[idx].n) fail(FMC_ERR_SIZE, "host buffer has %lld elements, vector has %lld", (long long)n, (long long)rt.nodes[idx].n);
    if (n > 0 && !h) fail(FMC_ERR_INVALID, "null host pointer");
    rt.materialize(idx);
    const float* src = rt.nodes[idx].buf;
    rt.staging.ensure(2 * kChunkElems * sizeof(float));
    rt.staging_quiesce();
    float* stage[2] = {(float*)rt.staging.host, (float*)rt.staging.host + kChunkElems};
    cudaEvent_t done[2] = {rt.ev_copy[0], rt.ev_copy[1]};
    // software pipeline: copy chunk c+1 while widening chunk c
    int64_t off = 0; int which = 0;
    int64_t pend_off[2] = {0, 0}, pend_m[2] = {0, 0};
    bool used[2] = {false, false};
    auto drain = [&](int w) {
        FMC_CUDA(cudaEventSynchronize(done[w]));
        const float* s = stage[w];
        T* d = h + pend_off[w];
        HostWorkers::get().parallel_for(pend_m[w], [&](int64_t b, int64_t e) {
            for (int64_t i = b; i < e; i++) d[i] = (T)s[i];                        // RandomVariableCuda.java:776-782
        });
        used[w] = false;
    };
    while (off < n) {
        const int64_t m = std::min<int64_t>(kChunkElems, n - off);
        if (used[which]) drain(which);
        FMC_CUDA(cudaMemcpyAsync(stage[which], src + off, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost, rt.stream));
        FMC_CUDA(cudaEventRecord(done[which], rt.stream));
        pend_off[which] = off; pend_m[which] = m; used[which] = true;
        which ^= 1; off += m;
    }
    if (used[which]) drain(which);
    if (used[which ^ 1]) drain(which ^ 1);
    rt.stats.d2h += sizeof(float) * (uint64_t)n;
}
void Runtime::download_f32(int32_t idx, float* h, int64_t n) { download_impl(*this, idx, h, n); }
void Runtime::download_f64(int32_t idx, double* h, int64_t n) { download_impl(*this, idx, h, n); }

}  // namespace fmc
