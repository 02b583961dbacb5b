// capi.cpp — the extern "C" surface declared in include/fmcuda.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>

#include "runtime.h"

using namespace fmc;

namespace {

template <typename F>
int guarded(F&& f) {
    Runtime& rt = Runtime::get();
    RuntimeLock lock(rt.mu);
    struct Held {                                    // lets Runtime::reduce drop the lock while it waits for its result
        Runtime& rt;
        Held(Runtime& r, RuntimeLock* l) : rt(r) { rt.held = l; }
        ~Held() { rt.held = nullptr; }
    } held(rt, &lock);
    try {
        // any thread may call: CUDA's current device is per thread, the runtime's device is not necessarily device 0
        static thread_local int bound_device = -1;
        if (rt.initialized && bound_device != rt.device) { FMC_CUDA(cudaSetDevice(rt.device)); bound_device = rt.device; }
        if (!rt.initialized) bound_device = -1;
        f(rt);
        return FMC_OK;
    } catch (const Fail& e) {
        return e.code;
    } catch (const std::bad_alloc&) {
        set_error("host out of memory");
        return FMC_ERR_OOM;
    } catch (const std::exception& e) {
        set_error("internal error: %s", e.what());
        return FMC_ERR_INVALID;
    }
}

void check_same_size(Runtime& rt, int32_t a, int32_t b) {
    if (rt.nodes[a].n != rt.nodes[b].n)
        fail(FMC_ERR_SIZE, "operand sizes differ: %lld vs %lld", (long long)rt.nodes[a].n, (long long)rt.nodes[b].n);
}

Operand N(int32_t idx) { return Operand{idx, 0.f}; }
Operand S(double s) { return Operand{-1, (float)s}; }     // (float)value: RandomVariableCuda.java:521, RVF:789

// a temporary created inside a compound op: owned only by its consumer
int32_t temp(Runtime& rt, int32_t idx) {
    Node& nd = rt.nodes[idx];
    nd.ext_refs = 0;
    rt.n_live_handles--;
    return idx;
}

void finish(Runtime& rt, int32_t idx, fmc_vec* out) {
    *out = rt.handle_of(idx);
    rt.auto_flush();
}

}  // namespace

extern "C" {

const char* fmc_last_error(void) { return get_error(); }

static void fmc_apply_env_options(Runtime& rt);
int fmc_init(int device_index) { return guarded([&](Runtime& rt) { const bool first = !rt.initialized; rt.init(device_index); if (first) fmc_apply_env_options(rt); }); }
int fmc_shutdown(void) { return guarded([&](Runtime& rt) { rt.shutdown(); }); }
int fmc_is_initialized(void) { return Runtime::get().initialized ? 1 : 0; }

int fmc_device_count(int* count) {
    return guarded([&](Runtime&) {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
        *count = c;
    });
}

int fmc_device_info(char* name, size_t name_len, int* sm_count, uint64_t* total_mem, int* cc_major, int* cc_minor) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        if (name && name_len) { std::strncpy(name, rt.prop.name, name_len - 1); name[name_len - 1] = 0; }
        if (sm_count) *sm_count = rt.prop.multiProcessorCount;
        if (total_mem) *total_mem = rt.prop.totalGlobalMem;
        if (cc_major) *cc_major = rt.prop.major;
        if (cc_minor) *cc_minor = rt.prop.minor;
    });
}

// ---- vectors ----
int fmc_vec_from_f64(const double* host, int64_t n, fmc_vec* out) {
    return guarded([&](Runtime& rt) { rt.require_init(); *out = rt.handle_of(rt.upload_f64(host, n)); });
}
int fmc_host_alloc(size_t bytes, void** out) { return guarded([&](Runtime& rt) { rt.require_init(); *out = rt.host_alloc(bytes); }); }
int fmc_host_free(void* p) { return guarded([&](Runtime& rt) { rt.require_init(); rt.host_free(p); }); }
int fmc_vec_from_f64_pinned(const double* pinned_host, int64_t n, fmc_vec* out) {
    return guarded([&](Runtime& rt) { rt.require_init(); *out = rt.handle_of(rt.upload_f64_pinned(pinned_host, n)); });
}
int fmc_vec_from_f32(const float* host, int64_t n, fmc_vec* out) {
    return guarded([&](Runtime& rt) { rt.require_init(); *out = rt.handle_of(rt.upload_f32(host, n)); });
}
int fmc_vec_alloc(int64_t n, fmc_vec* out) {
    return guarded([&](Runtime& rt) { rt.require_init(); *out = rt.handle_of(rt.new_leaf(n)); });
}
int fmc_vec_fill(double value, int64_t n, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        if (n < 0) fail(FMC_ERR_INVALID, "negative size");
        const int32_t r = rt.record(N_CONST, n, S(value));
        finish(rt, r, out);
    });
}
int fmc_vec_retain(fmc_vec v) { return guarded([&](Runtime& rt) { rt.retain(rt.resolve(v)); }); }
int fmc_vec_release(fmc_vec v) { return guarded([&](Runtime& rt) { rt.release_ext(rt.resolve(v)); }); }
int fmc_vec_size(fmc_vec v, int64_t* n) { return guarded([&](Runtime& rt) { *n = rt.nodes[rt.resolve(v)].n; }); }
int fmc_vec_to_f64(fmc_vec v, double* host, int64_t n) {
    return guarded([&](Runtime& rt) { rt.require_init(); rt.download_f64(rt.resolve(v), host, n); });
}
int fmc_vec_to_f32(fmc_vec v, float* host, int64_t n) {
    return guarded([&](Runtime& rt) { rt.require_init(); rt.download_f32(rt.resolve(v), host, n); });
}
int fmc_vec_get(fmc_vec v, int64_t i, double* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t idx = rt.resolve(v);
        if (i < 0 || i >= rt.nodes[idx].n) fail(FMC_ERR_SIZE, "index %lld out of range [0,%lld)", (long long)i, (long long)rt.nodes[idx].n);
        rt.materialize(idx);
        float f;
        FMC_CUDA(cudaMemcpyAsync(&f, rt.nodes[idx].buf + i, sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
        FMC_CUDA(cudaStreamSynchronize(rt.stream));
        *out = (double)f;
    });
}
int fmc_vec_device_ptr(fmc_vec v, void** device_ptr) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t idx = rt.resolve(v);
        rt.materialize(idx);
        FMC_CUDA(cudaStreamSynchronize(rt.stream));
        *device_ptr = rt.nodes[idx].buf;
    });
}

// ---- recorded elementwise operations ----
int fmc_op_vs(int opcode, fmc_vec a, double s, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a);
        const int64_t n = rt.nodes[x].n;
        int32_t r;
        switch (opcode) {
        case FMC_CAP:   r = rt.record(N_MIN, n, N(x), S(s)); break;       // RVF:757-760
        case FMC_FLOOR: r = rt.record(N_MAX, n, N(x), S(s)); break;       // RVF:772-775
        case FMC_ADD:   r = rt.record(N_ADD, n, N(x), S(s)); break;       // RVF:787-790
        case FMC_SUB:   r = rt.record(N_SUB, n, N(x), S(s)); break;       // RVF:802-805
        case FMC_BUS:   r = rt.record(N_SUB, n, S(s), N(x)); break;       // kernel busScalar: -a + b
        case FMC_MULT:  r = rt.record(N_MUL, n, N(x), S(s)); break;       // RVF:817-820
        case FMC_DIV:   r = rt.record(N_DIV, n, N(x), S(s)); break;       // RVF:832-835
        case FMC_VID:   r = rt.record(N_DIV, n, S(s), N(x)); break;       // kernel vidScalar: b / a
        case FMC_POW:   r = rt.record(N_POW, n, N(x), S(s)); break;       // RVF:847-850
        default: fail(FMC_ERR_INVALID, "fmc_op_vs: unknown opcode %d", opcode);
        }
        finish(rt, r, out);
    });
}

int fmc_op_v(int opcode, fmc_vec a, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a);
        const int64_t n = rt.nodes[x].n;
        int32_t r;
        switch (opcode) {
        case FMC_SQUARED: r = rt.record(N_MUL, n, N(x), N(x)); break;     // RVF:873-876, RVC:1290
        case FMC_SQRT:    r = rt.record(N_SQRT, n, N(x)); break;
        case FMC_EXP:     r = rt.record(N_EXP, n, N(x)); break;
        case FMC_LOG:     r = rt.record(N_LOG, n, N(x)); break;
        case FMC_SIN:     r = rt.record(N_SIN, n, N(x)); break;
        case FMC_COS:     r = rt.record(N_COS, n, N(x)); break;
        case FMC_INVERT:  r = rt.record(N_INV, n, N(x)); break;
        case FMC_ABS:     r = rt.record(N_ABS, n, N(x)); break;
        case FMC_ISNAN:   r = rt.record(N_ISNAN, n, N(x)); break;
        default: fail(FMC_ERR_INVALID, "fmc_op_v: unknown opcode %d", opcode);
        }
        finish(rt, r, out);
    });
}

int fmc_op_vv(int opcode, fmc_vec a, fmc_vec b, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a), y = rt.resolve(b);
        check_same_size(rt, x, y);
        const int64_t n = rt.nodes[x].n;
        int32_t r;
        switch (opcode) {
        case FMC_ADD:   r = rt.record(N_ADD, n, N(x), N(y)); break;       // RVF:981-984
        case FMC_SUB:   r = rt.record(N_SUB, n, N(x), N(y)); break;       // RVF:1011-1014
        case FMC_BUS:   r = rt.record(N_SUB, n, N(y), N(x)); break;       // RVF:1041-1044
        case FMC_MULT:  r = rt.record(N_MUL, n, N(x), N(y)); break;       // RVF:1073-1076
        case FMC_DIV:   r = rt.record(N_DIV, n, N(x), N(y)); break;       // RVF:1106-1109
        case FMC_VID:   r = rt.record(N_DIV, n, N(y), N(x)); break;       // RVF:1136-1139 (double division of floats == float division)
        case FMC_CAP:   r = rt.record(N_MIN, n, N(x), N(y)); break;       // RVF:1165-1168
        case FMC_FLOOR: r = rt.record(N_MAX, n, N(x), N(y)); break;       // RVF:1194-1197
        default: fail(FMC_ERR_INVALID, "fmc_op_vv: unknown opcode %d", opcode);
        }
        finish(rt, r, out);
    });
}

int fmc_op_vvs(int opcode, fmc_vec a, fmc_vec b, double s, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a), y = rt.resolve(b);
        check_same_size(rt, x, y);
        const int64_t n = rt.nodes[x].n;
        int32_t r;
        switch (opcode) {
        case FMC_ACCRUE: {        // x * (1.0f + y * p)      RVF:1222-1225, kernel.cu:224-231
            const int32_t t1 = temp(rt, rt.record(N_MUL, n, N(y), S(s)));
            const int32_t t2 = temp(rt, rt.record(N_ADD, n, N(t1), S(1.0)));
            r = rt.record(N_MUL, n, N(x), N(t2));
            break;
        }
        case FMC_DISCOUNT: {      // x / (1.0f + y * p)      RVF:1250-1253, kernel.cu:234-244
            const int32_t t1 = temp(rt, rt.record(N_MUL, n, N(y), S(s)));
            const int32_t t2 = temp(rt, rt.record(N_ADD, n, N(t1), S(1.0)));
            r = rt.record(N_DIV, n, N(x), N(t2));
            break;
        }
        case FMC_ADDPRODUCT: {    // x + y * s               RVF:1345-1348, kernel.cu:257-264
            const int32_t t1 = temp(rt, rt.record(N_MUL, n, N(y), S(s)));
            r = rt.record(N_ADD, n, N(x), N(t1));
            break;
        }
        default: fail(FMC_ERR_INVALID, "fmc_op_vvs: unknown opcode %d", opcode);
        }
        finish(rt, r, out);
    });
}

int fmc_op_vvv(int opcode, fmc_vec a, fmc_vec b, fmc_vec c, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a), y = rt.resolve(b), z = rt.resolve(c);
        check_same_size(rt, x, y);
        check_same_size(rt, x, z);
        const int64_t n = rt.nodes[x].n;
        int32_t r;
        switch (opcode) {
        case FMC_ADDPRODUCT: {    // x + y * z               RVF:1374-1377, kernel.cu:247-254
            const int32_t t1 = temp(rt, rt.record(N_MUL, n, N(y), N(z)));
            r = rt.record(N_ADD, n, N(x), N(t1));
            break;
        }
        case FMC_CHOOSE:          // x >= 0 ? y : z          RVF:1280-1282
            r = rt.record(N_CHOOSE, n, N(x), N(y), N(z));
            break;
        case FMC_ADDRATIO: {      // x + y / z               RVF:1409-1412
            const int32_t t1 = temp(rt, rt.record(N_DIV, n, N(y), N(z)));
            r = rt.record(N_ADD, n, N(x), N(t1));
            break;
        }
        case FMC_SUBRATIO: {      // x - y / z               RVF:1432-1435
            const int32_t t1 = temp(rt, rt.record(N_DIV, n, N(y), N(z)));
            r = rt.record(N_SUB, n, N(x), N(t1));
            break;
        }
        default: fail(FMC_ERR_INVALID, "fmc_op_vvv: unknown opcode %d", opcode);
        }
        finish(rt, r, out);
    });
}

int fmc_op_choose(fmc_vec trigger, fmc_vec if_nonneg, double s_nonneg, fmc_vec if_neg, double s_neg, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(trigger);
        const int64_t n = rt.nodes[x].n;
        Operand A = S(s_nonneg), B = S(s_neg);
        if (if_nonneg) { const int32_t y = rt.resolve(if_nonneg); check_same_size(rt, x, y); A = N(y); }
        if (if_neg) { const int32_t z = rt.resolve(if_neg); check_same_size(rt, x, z); B = N(z); }
        finish(rt, rt.record(N_CHOOSE, n, N(x), A, B), out);
    });
}

// ---- reductions ----
namespace {

// merge (count, value, M2) triples of the ranks in rank order (deterministic)
void merge_ranks(Runtime& rt, int mode, double part[3], bool already_global) {
    if (rt.comm_size <= 1 || already_global) return;             // in-kernel exchange: already the result of all ranks
    // Runtime::reduce left every rank's partial in h_result[4 * r + 0..2] (ncclAllGather behind the reduction kernel)
    const int R = rt.comm_size;
    (void)part;
    double c = 0.0, v = 0.0, m = 0.0;
    for (int r = 0; r < R; r++) {
        const double bc = rt.h_result[4 * r], bv = rt.h_result[4 * r + 1], bm = rt.h_result[4 * r + 2];
        if (bc == 0.0) continue;
        if (c == 0.0) { c = bc; v = bv; m = bm; continue; }
        const double tot = c + bc;
        if (mode == RM_MOMENTS) {
            const double delta = bv - v, w = bc / tot;
            m = m + bm + delta * delta * c * w;
            v = v + delta * w;
        } else if (mode == RM_MIN) v = (v != v || bv != bv) ? NAN : (v == 0.0 && bv == 0.0) ? ((std::signbit(v) || std::signbit(bv)) ? -0.0 : 0.0) : std::min(v, bv);   // java.lang.Math.min: -0 < +0
        else if (mode == RM_MAX) v = (v != v || bv != bv) ? NAN : (v == 0.0 && bv == 0.0) ? ((std::signbit(v) && std::signbit(bv)) ? -0.0 : 0.0) : std::max(v, bv);
        else v += bv;
        c = tot;
    }
    part[0] = c; part[1] = v; part[2] = m;
}

}  // namespace

int fmc_reduce(int kind, fmc_vec a, fmc_vec weights, double* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a);
        int32_t w = -1;
        if (kind == FMC_RED_AVERAGE_W || kind == FMC_RED_VARIANCE_W) {
            if (!weights) fail(FMC_ERR_INVALID, "weighted reduction needs a weight vector");
            w = rt.resolve(weights);
            check_same_size(rt, x, w);
        }
        ReduceSpec spec;
        double p[3];
        auto run = [&](const ReduceSpec& sp, int merge_mode) { const bool global = rt.reduce(x, sp, p); merge_ranks(rt, merge_mode, p, global); };
        // RVF:322-330 sums with Kahan's compensation: an infinite element makes the compensation term inf - inf = NaN, so every
        // element AFTER it turns the sum into NaN; only an infinity at the very last index survives as +-inf. The device sum
        // returns +-inf in both cases: when that happens (rare) the reference's answer is reconstructed with two more passes.
        auto kahan_infinity = [&](double sum) -> double {
            const int64_t n = rt.nodes[x].n;
            const int32_t d = temp(rt, rt.record(N_SUB, n, N(x), N(x)));          // inf - inf = NaN, finite - finite = 0
            const int32_t f = rt.record(N_ISNAN, n, N(d));
            ReduceSpec cs; cs.mode = RM_SUM;
            double q[3];
            const bool global = rt.reduce(f, cs, q);
            merge_ranks(rt, RM_SUM, q, global);
            rt.release_ext(f);
            if (q[1] != 1.0 || rt.comm_size > 1) return NAN;                     // (sharded: an infinity at the global last index is reported as NaN too)
            rt.materialize(x);
            float last = 0.f;
            FMC_CUDA(cudaMemcpyAsync(&last, rt.nodes[x].buf + (n - 1), sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
            FMC_CUDA(cudaStreamSynchronize(rt.stream));
            return std::isinf(last) ? sum : NAN;
        };
        switch (kind) {
        case FMC_RED_SUM:
        case FMC_RED_AVERAGE:
            spec.mode = RM_SUM;
            run(spec, RM_SUM);
            if (std::isinf(p[1])) p[1] = kahan_infinity(p[1]);
            if (kind == FMC_RED_SUM) *out = (p[0] == 0.0) ? 0.0 : p[1];
            else *out = (p[0] == 0.0) ? NAN : p[1] / p[0];                       // RVF:318-320, 333
            break;
        case FMC_RED_VARIANCE:
        case FMC_RED_SAMPLE_VARIANCE:
            spec.mode = RM_MOMENTS;
            run(spec, RM_MOMENTS);
            if (p[0] == 0.0) *out = NAN;                                         // RVF:364-366
            else if (p[0] == 1.0) *out = 0.0;                                    // RVF:361-363
            else {
                const double var = p[2] / p[0];                                  // RVF:381
                *out = (kind == FMC_RED_VARIANCE) ? var : var * p[0] / (p[0] - 1.0);   // RVF:418
            }
            break;
        case FMC_RED_MIN:
        case FMC_RED_MAX:
            spec.mode = (kind == FMC_RED_MIN) ? RM_MIN : RM_MAX;
            run(spec, spec.mode);
            if (p[0] == 0.0) *out = (kind == FMC_RED_MIN) ? 1.7976931348623157e308 : -1.7976931348623157e308;  // RVF:288, 303
            else *out = p[1];
            break;
        case FMC_RED_AVERAGE_W:
            spec.mode = RM_DOT; spec.weight = w;
            run(spec, RM_SUM);
            *out = (p[0] == 0.0) ? NAN : p[1] / p[0];                            // RVF:356
            break;
        case FMC_RED_VARIANCE_W: {
            rt.materialize(x);                                                   // two passes over x: keep it
            spec.mode = RM_DOT; spec.weight = w;
            run(spec, RM_SUM);
            if (p[0] == 0.0) { *out = NAN; break; }
            const double avg = p[1] / p[0];                                      // RVF:393
            ReduceSpec s2; s2.mode = RM_WSQ; s2.weight = w; s2.param = avg;
            run(s2, RM_SUM);
            *out = p[1];                                                         // RVF:406 (not divided by n)
            break;
        }
        default: fail(FMC_ERR_INVALID, "fmc_reduce: unknown kind %d", kind);
        }
    });
}

// ---- execution control ----
int fmc_flush(void) { return guarded([&](Runtime& rt) { rt.require_init(); rt.flush_all(); }); }
int fmc_sync(void) {
    return guarded([&](Runtime& rt) { rt.require_init(); rt.flush_all(); rt.sync_stream(); });
}
static void set_option_locked(Runtime& rt, const char* key, double value) {
    // everything that steers the code generator invalidates the cached tapes
    static const char* const keeps_cache[] = {"flush_threshold", "profile", "tape_upload_stream", "exchange", "exchange_timeout_s", "p2p_reduce", "zero_copy_reduce", "leaf_reduce_kernel", "batch_reduce", "regression_float_products"};
    bool keep = false;
    for (const char* k : keeps_cache) keep = keep || !std::strcmp(key, k);
    if (!keep) tape_cache_clear();
    if (!std::strcmp(key, "flush_threshold")) rt.opt.flush_threshold = (int64_t)value;
    else if (!std::strcmp(key, "tape_cache")) rt.opt.tape_cache = value != 0.0;
    else if (!std::strcmp(key, "max_regs")) rt.opt.max_regs = std::max(4, std::min((int)value, (int)TAPE_REGS));
    else if (!std::strcmp(key, "fuse")) { rt.opt.fuse = value != 0.0; if (rt.initialized) rt.flush_all(); }
    else if (!std::strcmp(key, "profile")) rt.opt.profile = value != 0.0;
    else if (!std::strcmp(key, "window_levels")) { rt.opt.window_levels = (int)value; rt.window_levels_now = 0; }
    else if (!std::strcmp(key, "window_elems")) rt.opt.window_elems = (int)value;
    else if (!std::strcmp(key, "window_cta_warps")) rt.opt.window_cta_warps = std::max(0, std::min((int)value, (int)TAPE_MAX_WARPS));
    else if (!std::strcmp(key, "window_reduce_min")) rt.opt.window_reduce_min = std::max(0, (int)value);
    else if (!std::strcmp(key, "window_ring_extra")) rt.opt.window_ring_extra = std::max(1, (int)value);
    else if (!std::strcmp(key, "ring_max")) rt.opt.ring_max = std::max(1, std::min((int)value, (int)TAPE_MAX_RING));
    else if (!std::strcmp(key, "ring_min")) rt.opt.ring_min = std::max(1, std::min((int)value, (int)TAPE_MAX_RING));
    else if (!std::strcmp(key, "target_ctas")) rt.opt.target_ctas = std::max(0, (int)value);
    else if (!std::strcmp(key, "tape_elems")) { const int e = (int)value; if (e != 0 && !tape_valid_elems(e)) fail(FMC_ERR_INVALID, "tape_elems must be 0 (auto), 4, 8 or 16"); rt.opt.tape_elems = e; }
    else if (!std::strcmp(key, "min_warps")) rt.opt.min_warps = std::max(1, (int)value);
    else if (!std::strcmp(key, "horizon")) rt.opt.horizon = std::max(0, (int)value);
    else if (!std::strcmp(key, "pipeline")) rt.opt.pipeline = value != 0.0;
    else if (!std::strcmp(key, "max_sets")) rt.opt.max_sets = std::max(1, std::min((int)value, 4));
    else if (!std::strcmp(key, "grid_limit")) rt.opt.grid_limit = std::max(0, (int)value);
    else if (!std::strcmp(key, "brownian_blocks_per_sm")) rt.opt.brownian_blocks_per_sm = std::max(0, (int)value);
    else if (!std::strcmp(key, "fuse_ops")) rt.opt.fuse_ops = value != 0.0;
    else if (!std::strcmp(key, "fuse_ops2")) rt.opt.fuse_ops2 = value != 0.0;
    else if (!std::strcmp(key, "regression_float_products")) rt.opt.regression_float_products = value != 0.0;
    else if (!std::strcmp(key, "batch_reduce")) { rt.opt.batch_reduce = value != 0.0; rt.flush_batches.clear(); rt.prefetched.clear(); }
    else if (!std::strcmp(key, "tape_upload_stream")) rt.opt.tape_upload_stream = value != 0.0;
    else if (!std::strcmp(key, "p2p_reduce")) rt.opt.p2p_reduce = value != 0.0;
    else if (!std::strcmp(key, "exchange")) { const int m = (int)value; if (m < 0 || m > 2) fail(FMC_ERR_INVALID, "exchange must be 0 (NCCL), 1 (in-kernel peer memory) or 2 (shared host memory)"); rt.opt.exchange = m; }
    else if (!std::strcmp(key, "exchange_timeout_s")) rt.opt.exchange_timeout_s = std::max(1.0, value);
    else if (!std::strcmp(key, "zero_copy_reduce")) rt.opt.zero_copy_reduce = value != 0.0;
    else if (!std::strcmp(key, "leaf_reduce_kernel")) rt.opt.leaf_reduce_kernel = value != 0.0;
    else if (!std::strcmp(key, "cta_warps")) rt.opt.cta_warps = std::max(1, std::min((int)value, (int)TAPE_MAX_WARPS));
    else fail(FMC_ERR_INVALID, "unknown option '%s'", key);
}
// FMC_OPTIONS="key=value,key=value": applied once at fmc_init (tuning experiments without touching the caller)
static void fmc_apply_env_options(Runtime& rt) {
    const char* env = std::getenv("FMC_OPTIONS");
    if (!env) return;
    std::string s(env);
    size_t at = 0;
    while (at < s.size()) {
        size_t end = s.find(',', at);
        if (end == std::string::npos) end = s.size();
        const std::string kv = s.substr(at, end - at);
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) set_option_locked(rt, kv.substr(0, eq).c_str(), std::atof(kv.c_str() + eq + 1));
        at = end + 1;
    }
}
int fmc_set_option(const char* key, double value) {
    return guarded([&](Runtime& rt) { set_option_locked(rt, key, value); });
}
int fmc_get_option(const char* key, double* value) {
    return guarded([&](Runtime& rt) {
        uint64_t c_hits = 0, c_misses = 0, c_entries = 0;
        tape_cache_stats(&c_hits, &c_misses, &c_entries);
        if (!std::strcmp(key, "flush_threshold")) *value = (double)rt.opt.flush_threshold;
        else if (!std::strcmp(key, "tape_cache")) *value = rt.opt.tape_cache ? 1.0 : 0.0;
        else if (!std::strcmp(key, "max_regs")) *value = rt.opt.max_regs;
        else if (!std::strcmp(key, "tape_cache_hits")) *value = (double)c_hits;
        else if (!std::strcmp(key, "tape_cache_misses")) *value = (double)c_misses;
        else if (!std::strcmp(key, "tape_cache_entries")) *value = (double)c_entries;
        else if (!std::strcmp(key, "fuse")) *value = rt.opt.fuse ? 1.0 : 0.0;
        else if (!std::strcmp(key, "profile")) *value = rt.opt.profile ? 1.0 : 0.0;
        else if (!std::strcmp(key, "window_levels")) *value = (double)rt.opt.window_levels;
        else if (!std::strcmp(key, "ring_max")) *value = rt.opt.ring_max;
        else if (!std::strcmp(key, "ring_min")) *value = rt.opt.ring_min;
        else if (!std::strcmp(key, "target_ctas")) *value = rt.opt.target_ctas;
        else if (!std::strcmp(key, "tape_elems")) *value = rt.opt.tape_elems;
        else if (!std::strcmp(key, "min_warps")) *value = rt.opt.min_warps;
        else if (!std::strcmp(key, "horizon")) *value = rt.opt.horizon;
        else if (!std::strcmp(key, "pipeline")) *value = rt.opt.pipeline ? 1.0 : 0.0;
        else if (!std::strcmp(key, "max_sets")) *value = rt.opt.max_sets;
        else if (!std::strcmp(key, "grid_limit")) *value = rt.opt.grid_limit;
        else if (!std::strcmp(key, "brownian_blocks_per_sm")) *value = rt.opt.brownian_blocks_per_sm;
        else if (!std::strcmp(key, "fuse_ops")) *value = rt.opt.fuse_ops ? 1.0 : 0.0;
        else if (!std::strcmp(key, "fuse_ops2")) *value = rt.opt.fuse_ops2 ? 1.0 : 0.0;
        else if (!std::strcmp(key, "regression_float_products")) *value = rt.opt.regression_float_products ? 1.0 : 0.0;
        else if (!std::strcmp(key, "batch_reduce")) *value = rt.opt.batch_reduce ? 1.0 : 0.0;
        else if (!std::strcmp(key, "p2p_reduce")) *value = rt.opt.p2p_reduce ? 1.0 : 0.0;
        else if (!std::strcmp(key, "p2p_ready")) *value = rt.p2p_ready ? 1.0 : 0.0;
        else if (!std::strcmp(key, "exchange")) *value = rt.opt.exchange;
        else if (!std::strcmp(key, "exchange_timeout_s")) *value = rt.opt.exchange_timeout_s;
        else if (!std::strcmp(key, "exchange_in_use")) *value = rt.use_xhost() ? 2.0 : rt.use_p2p() ? 1.0 : 0.0;
        else if (!std::strcmp(key, "device_index")) *value = rt.device;
        else if (!std::strcmp(key, "zero_copy_reduce")) *value = rt.opt.zero_copy_reduce ? 1.0 : 0.0;
        else if (!std::strcmp(key, "leaf_reduce_kernel")) *value = rt.opt.leaf_reduce_kernel ? 1.0 : 0.0;
        else if (!std::strcmp(key, "cta_warps")) *value = rt.opt.cta_warps;
        else if (!std::strcmp(key, "profile_touched_bytes")) *value = (double)rt.prof_touched;
        else if (!std::strcmp(key, "host_us_codegen")) *value = rt.hostprof.codegen;
        else if (!std::strcmp(key, "host_us_launch")) *value = rt.hostprof.launch;
        else if (!std::strcmp(key, "host_us_sync")) *value = rt.hostprof.sync;
        else if (!std::strcmp(key, "host_us_upload")) *value = rt.hostprof.upload;
        else if (!std::strcmp(key, "host_us_upload_wait")) *value = rt.hostprof.upload_wait;
        else fail(FMC_ERR_INVALID, "unknown option '%s'", key);
    });
}

int fmc_get_stats(fmc_stats* out) {
    return guarded([&](Runtime& rt) {
        std::memset(out, 0, sizeof(*out));
        out->bytes_in_use = rt.pool.bytes_in_use; out->bytes_cached = rt.pool.bytes_cached;
        out->bytes_reserved = rt.pool.bytes_reserved; out->bytes_high_water = rt.pool.high_water;
        out->n_alloc = rt.pool.n_alloc; out->n_alloc_reused = rt.pool.n_reused;
        out->n_ops_recorded = rt.stats.n_ops; out->n_kernels = rt.stats.n_kernels;
        out->n_tape_kernels = rt.stats.n_tape_kernels; out->n_tape_instr = rt.stats.n_tape_instr;
        out->n_nodes_stored = rt.stats.n_stored; out->n_nodes_fused = rt.stats.n_fused; out->n_flushes = rt.stats.n_flushes;
        out->h2d_bytes = rt.stats.h2d; out->d2h_bytes = rt.stats.d2h;
        out->live_handles = (uint64_t)rt.n_live_handles; out->pending_nodes = (uint64_t)rt.n_lazy;
    });
}
int fmc_reset_stats(void) {
    return guarded([&](Runtime& rt) { rt.stats = Stats{}; rt.hostprof = HostProfile{}; rt.pool.n_alloc = rt.pool.n_reused = 0; rt.pool.high_water = rt.pool.bytes_in_use; });
}
int fmc_pool_trim(void) {
    return guarded([&](Runtime& rt) { rt.require_init(); FMC_CUDA(cudaStreamSynchronize(rt.stream)); rt.pool.trim(); });
}
int fmc_pool_purge(void) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        FMC_CUDA(cudaStreamSynchronize(rt.stream));
        brownian_release_caches(rt);
        rt.pool.purge();
    });
}

int fmc_profile_read(double* tape_ms, uint64_t* tape_algorithmic_bytes, uint64_t* tape_launches) {
    return guarded([&](Runtime& rt) { rt.require_init(); rt.profile_read(tape_ms, tape_algorithmic_bytes, tape_launches); });
}

int fmc_timer_start(void) { return guarded([&](Runtime& rt) { rt.require_init(); FMC_CUDA(cudaEventRecord(rt.ev_start, rt.stream)); }); }
int fmc_timer_stop(float* ms) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        FMC_CUDA(cudaEventRecord(rt.ev_stop, rt.stream));
        FMC_CUDA(cudaEventSynchronize(rt.ev_stop));
        FMC_CUDA(cudaEventElapsedTime(ms, rt.ev_start, rt.ev_stop));
    });
}

// ---- regression ----
int fmc_regression_normal_eq(const fmc_vec* basis, const double* scalars, int k, fmc_vec y, double* XtX, double* Xty) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        if (k < 1 || k > REG_MAX_K) fail(FMC_ERR_INVALID, "regression: k=%d outside [1,%d]", k, REG_MAX_K);
        const int32_t yi = rt.resolve(y);
        const int64_t n = rt.nodes[yi].n;
        RegressionParams P{};
        std::vector<int32_t> need;
        need.push_back(yi);
        int32_t bidx[REG_MAX_K];
        for (int i = 0; i < k; i++) {
            bidx[i] = -1;
            if (basis[i]) {
                bidx[i] = rt.resolve(basis[i]);
                check_same_size(rt, yi, bidx[i]);
                need.push_back(bidx[i]);
            }
        }
        std::vector<int32_t> lazy;
        for (int32_t v : need) if (rt.nodes[v].state == NS_LAZY) lazy.push_back(v);
        if (!lazy.empty()) rt.run_cone(lazy, nullptr);
        const int m = k * (k + 1) / 2 + k;
        auto all_nan = [&] { for (int i = 0; i < k * k; i++) XtX[i] = NAN; for (int i = 0; i < k; i++) Xty[i] = NAN; };
        if (n == 0 && rt.comm_size == 1) { all_nan(); return; }
        P.n = n; P.k = k; P.y = rt.nodes[yi].buf;
        for (int i = 0; i < k; i++) {
            P.basis[i] = bidx[i] >= 0 ? rt.nodes[bidx[i]].buf : nullptr;
            P.scalars[i] = bidx[i] >= 0 ? 0.f : (float)scalars[i];
        }
        P.partials = rt.d_partials; P.counter = rt.d_counter + 1; P.result = rt.d_result + 64;
        P.float_products = rt.opt.regression_float_products ? 1 : 0;
        static int per_sm_k[REG_MAX_K + 1] = {0};
        if (!per_sm_k[k]) per_sm_k[k] = regression_max_blocks_per_sm(k);
        const int per_sm = per_sm_k[k];
        const int64_t tiles = (n + regression_tile_elems() - 1) / regression_tile_elems();
        int grid = (int)std::min<int64_t>(tiles, (int64_t)per_sm * rt.sm_count);
        grid = std::max(1, std::min(grid, rt.max_grid));
        // an empty slice of a sharded vector contributes zeros and still joins the all-reduce (the other ranks are in it)
        if (n > 0) { FMC_CUDA(launch_regression(P, grid, rt.stream)); rt.stats.n_kernels++; }
        else FMC_CUDA(cudaMemsetAsync(rt.d_result + 64, 0, sizeof(double) * (size_t)m, rt.stream));
        double cnt = (double)n;
        if (rt.comm_size > 1) {
            // one all-reduce of the k(k+1)/2 + k sums plus the path count
            FMC_CUDA(cudaMemcpyAsync(rt.d_result + 64 + m, &cnt, sizeof(double), cudaMemcpyHostToDevice, rt.stream));
            rt.allreduce_sum(rt.d_result + 64, m + 1);
        }
        FMC_CUDA(cudaMemcpyAsync(rt.h_result + 64, rt.d_result + 64, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, rt.stream));
        FMC_CUDA(cudaStreamSynchronize(rt.stream));
        if (rt.comm_size > 1) cnt = rt.h_result[64 + m];
        if (cnt == 0.0) { all_nan(); return; }
        const double* s = rt.h_result + 64;
        int t = 0;
        for (int i = 0; i < k; i++)
            for (int j = i; j < k; j++, t++) {
                double v;
                if (bidx[i] < 0 && bidx[j] < 0) v = scalars[i] * scalars[j];      // deterministic x deterministic: double (RVF:1059-1061)
                else v = s[t] / cnt;
                XtX[i * k + j] = v; XtX[j * k + i] = v;
            }
        for (int i = 0; i < k; i++, t++) Xty[i] = s[t] / cnt;
    });
}

#ifdef FMC_TAPE_TIMING
int fmc_debug_stamps(unsigned long long* out) {
    return guarded([&](Runtime& rt) { rt.require_init(); FMC_CUDA(cudaStreamSynchronize(rt.stream)); FMC_CUDA(fmc::tape_read_stamps_e16(out)); });
}
#endif

// ---- Brownian ----
int fmc_brownian_generate(int seed_mode, int64_t seed, int T, int F, int64_t p0, int64_t p1, const double* sqrt_dt, fmc_vec* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        if (T < 1 || F < 1 || p0 < 0 || p1 < p0 || !sqrt_dt || !out) fail(FMC_ERR_INVALID, "fmc_brownian_generate: bad arguments");
        std::vector<int32_t> nodes((size_t)T * F);
        brownian_generate(rt, seed_mode, seed, T, F, p0, p1, sqrt_dt, nodes.data());
        for (size_t i = 0; i < nodes.size(); i++) out[i] = rt.handle_of(nodes[i]);
    });
}
int fmc_mt19937_raw(int seed_mode, int64_t seed, uint64_t skip, int64_t count, uint32_t* host_out) {
    return guarded([&](Runtime& rt) { rt.require_init(); mt19937_raw(rt, seed_mode, seed, skip, count, host_out); });
}

// ---- order statistics (RVF:472-602): radix select on the device (order_kernel.cu), also for sharded vectors ----
namespace {
int64_t quantile_index(int64_t n, double q) {                                    // RVF:484
    long long idx = (long long)std::floor((double)(n + 1) * q - 1.0 + 0.5);
    return std::min<long long>(std::max<long long>(idx, 0), n - 1);
}
uint32_t sort_key_host(float f) {
    if (f != f) return 0xffffffffu;
    uint32_t u; std::memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
float key_to_float(uint32_t key) {
    if (key == 0xffffffffu) return NAN;
    const uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    float f; std::memcpy(&f, &u, 4);
    return f;
}
int order_grid(Runtime& rt, int64_t n) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 4095) / 4096, (int64_t)rt.sm_count * 8));
}
// device scratch d_result[512 .. 512+count) zeroed, kernel by `launch`, summed over the ranks, copied to h_result[512..]
const double* order_pass(Runtime& rt, int count, const std::function<void(double*)>& launch) {
    double* d = rt.d_result + 512;
    FMC_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)count, rt.stream));
    launch(d);
    rt.stats.n_kernels++;
    if (rt.comm_size > 1) rt.allreduce_sum(d, count);
    FMC_CUDA(cudaMemcpyAsync(rt.h_result + 512, d, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    rt.stats.d2h += sizeof(double) * (uint64_t)count;
    return rt.h_result + 512;
}
int64_t global_size(Runtime& rt, int64_t local_n) {
    if (rt.comm_size <= 1) return local_n;
    double cnt = (double)local_n;
    FMC_CUDA(cudaMemcpyAsync(rt.d_result + 200, &cnt, sizeof(double), cudaMemcpyHostToDevice, rt.stream));
    rt.allreduce_sum(rt.d_result + 200, 1);
    FMC_CUDA(cudaMemcpyAsync(rt.h_result + 200, rt.d_result + 200, sizeof(double), cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    return (int64_t)rt.h_result[200];
}
// key of the element with (global) rank r in Arrays.sort order; optionally the number of elements below that key
uint32_t select_key(Runtime& rt, const float* x, int64_t n, int64_t r) {
    uint32_t prefix = 0u, mask = 0u;
    for (int shift = 24; shift >= 0; shift -= 8) {
        const double* h = order_pass(rt, 256, [&](double* d) { FMC_CUDA(launch_select_hist(x, n, prefix, mask, shift, d, order_grid(rt, n), rt.stream)); });
        int64_t cum = 0; int digit = 255;
        for (int b = 0; b < 256; b++) {
            const int64_t c = (int64_t)h[b];
            if (r < cum + c) { digit = b; break; }
            cum += c;
        }
        r -= cum;
        prefix |= (uint32_t)digit << shift;
        mask |= 0xffu << shift;
    }
    return prefix;
}
}  // namespace

int fmc_quantile(fmc_vec a, double q, double* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a);
        rt.materialize(x);
        const int64_t n = rt.nodes[x].n, N = global_size(rt, n);
        if (N == 0) { *out = NAN; return; }
        *out = (double)key_to_float(select_key(rt, rt.nodes[x].buf, n, quantile_index(N, q)));
    });
}
int fmc_quantile_expectation(fmc_vec a, double q0, double q1, double* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        const int32_t x = rt.resolve(a);
        rt.materialize(x);
        const int64_t n = rt.nodes[x].n, N = global_size(rt, n);
        if (N == 0) { *out = NAN; return; }
        if (q0 > q1) std::swap(q0, q1);                                          // RVF:509-511
        const int64_t i0 = quantile_index(N, q0), i1 = quantile_index(N, q1);
        const float* buf = rt.nodes[x].buf;
        const uint32_t k0 = select_key(rt, buf, n, i0), k1 = select_key(rt, buf, n, i1);
        // mean of sorted[i0 .. i1] (RVF:519-523): everything strictly between the two values plus the copies of the boundary
        // values whose ranks fall inside the range
        const double* s = order_pass(rt, 5, [&](double* d) { FMC_CUDA(launch_range_stats(buf, n, k0, k1, d, order_grid(rt, n), rt.stream)); });
        const double v0 = (double)key_to_float(k0), v1 = (double)key_to_float(k1);
        const int64_t lt0 = (int64_t)s[0], eq0 = (int64_t)s[1], mid = (int64_t)s[2], eq1 = (int64_t)s[4];
        double e;
        if (k0 == k1) e = v0 * (double)(i1 - i0 + 1);
        else {
            const int64_t n0 = lt0 + eq0 - i0;                                   // copies of v0 with rank >= i0
            const int64_t n1 = i1 - (lt0 + eq0 + mid) + 1;                       // copies of v1 with rank <= i1
            (void)eq1;
            e = v0 * (double)n0 + s[3] + v1 * (double)n1;
        }
        *out = e / (double)(i1 - i0 + 1);
    });
}
int fmc_histogram(fmc_vec a, const double* pts, int m, double* out) {
    return guarded([&](Runtime& rt) {
        rt.require_init();
        if (m < 0 || m > 100) fail(FMC_ERR_INVALID, "histogram: %d interval points (at most 100)", m);
        const int32_t x = rt.resolve(a);
        rt.materialize(x);
        const int64_t n = rt.nodes[x].n, N = global_size(rt, n);
        if (m > 0) {
            FMC_CUDA(cudaMemcpyAsync(rt.d_result + 256, pts, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, rt.stream));
            FMC_CUDA(cudaStreamSynchronize(rt.stream));                          // pageable source
        }
        const double* c = order_pass(rt, m + 1, [&](double* d) {
            if (n > 0) FMC_CUDA(launch_histogram(rt.nodes[x].buf, n, rt.d_result + 256, m, d, order_grid(rt, n), rt.stream));
        });
        for (int k = 0; k <= m; k++) out[k] = (N > 0) ? c[k] / (double)N : c[k];  // RVF:571-600
    });
}

// ---- multi GPU ----
int fmc_comm_get_unique_id(char* id) { return guarded([&](Runtime&) { comm_get_unique_id(id); }); }
int fmc_comm_init(int rank, int nranks, const char* id) { return guarded([&](Runtime& rt) { rt.require_init(); comm_init(rt, rank, nranks, id); }); }
int fmc_comm_destroy(void) { return guarded([&](Runtime& rt) { comm_destroy(rt); }); }
int fmc_comm_info(int* rank, int* nranks) {
    return guarded([&](Runtime& rt) { if (rank) *rank = rt.comm_rank; if (nranks) *nranks = rt.comm_size; });
}

}  // extern "C"
