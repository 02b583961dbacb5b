// fmcuda_emu.cpp — TEST INFRASTRUCTURE, never shipped and never loaded by the product.
//
// An instruction-level emulator of the op-tape ISA (finmath-lib-cuda-extensions_b200/csrc/tape_isa.h) plus host
// stand-ins for the handful of CUDA runtime calls the C++ runtime makes. Linking the product's own host sources
// (runtime.cpp, pool.cpp, codegen.cpp, capi.cpp, ...) against this file instead of libcudart + the sm_100a kernels
// yields tests/emu/libfmcuda_emu.so, which exports the same C ABI. It exists for ONE purpose: to check, on a machine
// without a GPU, that the code generator emits well-formed tapes — every TMA ring slot is armed, waited for and
// re-armed in a legal order, nothing is read back through the ring that the same launch wrote, cross-chunk
// prefetches (prologue / T_LOADN) line up with the chunk loop, register-file slots are written before they are
// read — and that those tapes compute the oracle's values under the documented semantics of each opcode.
// It is NOT a fallback: the product library has no code path into it, and the GPU tests never load it.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "kernels.h"
#include "tape_isa.h"

// ---------------------------------------------------------------------------------------------------------------
// CUDA runtime stand-ins: "device" memory is host memory, streams are synchronous
// ---------------------------------------------------------------------------------------------------------------
extern "C" {
cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "tape-ISA emulator (no GPU)");
    p->multiProcessorCount = 148; p->major = 10; p->minor = 0;
    p->sharedMemPerMultiprocessor = 233472; p->totalGlobalMem = 8ull << 30;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 512, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = 4ull << 30; *t = 8ull << 30; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)0x1; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)0x1; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)0x1; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulator: tape check failed (see stderr)"; }
}

namespace fmc {

namespace {

std::string g_emu_error;

struct Check { std::string msg; };
[[noreturn]] void bad(const char* fmt, int a = 0, int b = 0, int c = 0) {
    char buf[256]; std::snprintf(buf, sizeof(buf), fmt, a, b, c);
    throw Check{buf};
}

float jminf(float a, float b) {      // java.lang.Math.min: NaN propagating, -0 < +0
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.f && b == 0.f) return (std::signbit(a) || std::signbit(b)) ? -0.f : 0.f;
    return a < b ? a : b;
}
float jmaxf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.f && b == 0.f) return (std::signbit(a) && std::signbit(b)) ? -0.f : 0.f;
    return a > b ? a : b;
}
float jpow(float x, float e) {       // mirrors f_pow in tape_interp.cuh
    const double dx = (double)x, de = (double)e;
    if (de != de) return (float)de;
    if (de == 0.0) return 1.0f;
    if (dx != dx) return x;
    if (std::isinf(de) && std::fabs(dx) == 1.0) return NAN;
    if (de == 2.0) return x * x;
    if (de == 1.0) return x;
    return (float)std::pow(dx, de);
}

constexpr int CMAX = 32 * TAPE_E_MAX;      // largest chunk; the geometry of a launch is P.elems

struct Warp {
    const TapeParams& P;
    const int C;                              // paths per chunk of this launch
    std::vector<float> slots;                 // [n_slots][C]
    std::vector<int> issued, waited;          // per ring slot: TMA copies armed / waited for
    std::vector<char> reg_written;            // per slot: register-file slot holds a value of the current chunk
    std::vector<long long> slot_chunk;        // chunk the ring slot's (pending or landed) data belongs to
    std::set<const float*>& stored;           // buffers written by this launch
    explicit Warp(const TapeParams& p, std::set<const float*>& st)
        : P(p), C(tape_chunk(p.elems)), slots((size_t)p.n_slots * tape_chunk(p.elems), NAN), issued(p.n_ring, 0), waited(p.n_ring, 0), reg_written(p.n_slots, 0),
          slot_chunk(p.n_ring, -1), stored(st) {}

    void load(uint32_t slot, uint32_t pidx, long long chunk) {
        if ((int)slot >= P.n_ring) bad("T_LOAD into slot %d outside the ring (n_ring %d)", (int)slot, P.n_ring);
        if ((int)pidx >= P.n_ptrs) bad("pointer index %d out of range", (int)pidx);
        if (issued[slot] != waited[slot]) bad("ring slot %d re-armed before its previous copy was waited for", (int)slot);
        const float* src = P.ptrs[pidx];
        if (stored.count(src)) bad("TMA read-back of a buffer (ptr %d) written by the same launch", (int)pidx);
        const long long base = chunk * C;
        if (base >= P.n) bad("TMA copy of a chunk past the end of the vector");
        const long long m = std::min<long long>(C, ((P.n - base) + 3) / 4 * 4);
        std::memcpy(&slots[(size_t)slot * C], src + base, sizeof(float) * (size_t)std::min<long long>(m, P.n - base));
        issued[slot]++;
        slot_chunk[slot] = chunk;
    }
    void wait(uint32_t slot) {
        if ((int)slot >= P.n_ring) bad("wait on slot %d outside the ring", (int)slot);
        if (issued[slot] != waited[slot] + 1) bad("wait on ring slot %d with no copy in flight", (int)slot);
        waited[slot]++;
    }
    const float* read(uint32_t slot, long long chunk) {
        if ((int)slot >= P.n_slots) bad("slot %d out of range (n_slots %d)", (int)slot, P.n_slots);
        if ((int)slot < P.n_ring) {
            if (issued[slot] != waited[slot]) bad("ring slot %d read while its copy is still in flight (missing _W)", (int)slot);
            if (slot_chunk[slot] != chunk) bad("ring slot %d holds data of another chunk", (int)slot);
        } else if (!reg_written[slot]) bad("register-file slot %d read before it was written in this chunk", (int)slot);
        return &slots[(size_t)slot * C];
    }
};

struct Partial { double c = 0, s = 0, s2 = 0, mn = 0, mx = 0; };

void run_warp(const TapeParams& P, long long first_chunk, long long stride, std::set<const float*>& stored, Partial& part,
              std::vector<double>& values /* RM_MOMENTS two-pass */) {
    const int C = tape_chunk(P.elems);
    const int SHIFT = tape_slot_shift(P.elems);
    const long long n_chunks = (P.n + C - 1) / C;
    std::vector<Warp> sets;
    for (int u = 0; u < P.n_sets; u++) sets.emplace_back(P, stored);
    if (P.instr[P.n_prologue].x != T_END) bad("prologue is not closed by T_END");
    // prologue: once per slot set, for the warp's first n_sets chunks
    for (int u = 0; u < P.n_sets; u++) {
        const long long chunk = first_chunk + u * stride;
        if (chunk >= n_chunks) continue;
        for (int pc = 0; pc < P.n_prologue; pc++) {
            const TapeInstr in = P.instr[pc];
            if ((in.x & ((1u << SHIFT) - 1u)) != T_LOAD) bad("prologue instruction %d is not a T_LOAD", pc);
            sets[u].load(in.x >> SHIFT, in.y, chunk);
        }
    }
    long long k = 0;
    for (long long chunk = first_chunk; chunk < n_chunks; chunk += stride, k++) {
        Warp& w = sets[(size_t)(k % P.n_sets)];
        const long long base = chunk * C;
        const long long next = chunk + stride * P.n_sets;
        float acc[CMAX]; bool pred[CMAX];
        // the kernel does not clear acc between chunks: a tape that read it before defining it would see the previous chunk's
        // values. Poisoned here so that such a tape fails the value comparison.
        for (int e = 0; e < C; e++) { acc[e] = std::nanf(""); pred[e] = false; }
        std::fill(w.reg_written.begin(), w.reg_written.end(), 0);
        const float* endb = nullptr;
        for (int pc = P.n_prologue + 1;; pc++) {
            if (pc >= P.n_instr) bad("ran off the end of the tape");
            const TapeInstr in = P.instr[pc];
            const uint32_t op = in.x & ((1u << SHIFT) - 1u), slot = in.x >> SHIFT;
            float imm; std::memcpy(&imm, &in.y, 4);
            if (op == T_END) {
                if (in.y != 0u) {
                    if ((int)slot < P.n_ring) bad("T_END reads a ring slot");
                    endb = w.read(slot, chunk);
                }
                break;
            }
            if (op == T_ADDAFF_S || op == T_ADDAFF_W) {
                if (pc + 1 >= P.n_instr) bad("two-word instruction at the end of the tape");
                if (P.instr[pc + 1].x != T_END) bad("the word behind two-word opcode %d is not its extension word", (int)op);
                float imm2; std::memcpy(&imm2, &P.instr[pc + 1].y, 4);
                if (op == T_ADDAFF_W) w.wait(slot);
                const float* b = w.read(slot, chunk);
                for (int e = 0; e < C; e++) { float t = b[e] + imm; t = t * imm2; acc[e] = acc[e] + t; }
                pc++;
                continue;
            }
            if (op == T_MULADDMUL || op == T_RATIO || op == T_ADDAFFDISC_S || op == T_ADDAFFDISC_W || op == T_ADDAFFDISC_SL || op == T_ADDAFFDISC_WL
                || op == T_RATIOACC_S || op == T_RATIOACC_W || op == T_AXPYST_S || op == T_RATIOACC_A) {
                const int ext = (op == T_AXPYST_S) ? 5 : (op == T_RATIO || op == T_ADDAFFDISC_SL || op == T_ADDAFFDISC_WL || op == T_RATIOACC_S || op == T_RATIOACC_W || op == T_RATIOACC_A) ? 3 : 2;
                if (pc + ext >= P.n_instr) bad("multi-word instruction at the end of the tape");
                float im[6] = {imm, 0.f, 0.f, 0.f, 0.f, 0.f};
                uint32_t ey[6] = {in.y, 0, 0, 0, 0, 0}, ex[6] = {0, 0, 0, 0, 0, 0};
                for (int k = 1; k <= ext; k++) {
                    const bool names_slot = (k == 3) && (op == T_RATIOACC_S || op == T_RATIOACC_W || op == T_AXPYST_S || op == T_RATIOACC_A);
                    if ((P.instr[pc + k].x & ((1u << SHIFT) - 1u)) != T_END) bad("extension word %d of opcode %d carries an opcode", k, (int)op);
                    if (!names_slot && P.instr[pc + k].x != T_END) bad("extension word %d of opcode %d is not a plain T_END word", k, (int)op);
                    std::memcpy(&im[k], &P.instr[pc + k].y, 4);
                    ey[k] = P.instr[pc + k].y; ex[k] = P.instr[pc + k].x >> SHIFT;
                }
                if (op == T_MULADDMUL) {
                    for (int e = 0; e < C; e++) { float t = acc[e] * im[0]; t = t + im[1]; acc[e] = t * im[2]; }
                } else if (op == T_RATIO) {
                    for (int e = 0; e < C; e++) { float t = acc[e] * im[0]; t = t + im[1]; t = im[2] / t; acc[e] = t * im[3]; }
                } else if (op == T_RATIOACC_S || op == T_RATIOACC_W) {
                    if (op == T_RATIOACC_W) w.wait(slot);
                    const float* b = w.read(slot, chunk);
                    const uint32_t s2 = ex[3];
                    if ((int)s2 < P.n_ring) bad("T_RATIOACC accumulates into a ring slot");
                    for (int e = 0; e < C; e++) { float t = b[e] * im[0]; t = t + im[1]; t = im[2] / t; acc[e] = t * im[3]; }
                    const float* c2 = w.read(s2, chunk);
                    for (int e = 0; e < C; e++) acc[e] = acc[e] + c2[e];
                    std::memcpy(&w.slots[(size_t)s2 * C], acc, sizeof(float) * (size_t)C);
                } else if (op == T_RATIOACC_A) {
                    const uint32_t s2 = ex[3];
                    if ((int)slot < P.n_ring || (int)slot >= P.n_slots) bad("T_RATIOACC_A parks acc in slot %d outside the register file", (int)slot);
                    if ((int)s2 < P.n_ring) bad("T_RATIOACC_A accumulates into a ring slot");
                    if (s2 == slot) bad("T_RATIOACC_A with both operands in one slot");
                    std::memcpy(&w.slots[(size_t)slot * C], acc, sizeof(float) * (size_t)C);
                    w.reg_written[slot] = 1;
                    for (int e = 0; e < C; e++) { float t = acc[e] * im[0]; t = t + im[1]; t = im[2] / t; acc[e] = t * im[3]; }
                    const float* c2 = w.read(s2, chunk);
                    for (int e = 0; e < C; e++) acc[e] = acc[e] + c2[e];
                    std::memcpy(&w.slots[(size_t)s2 * C], acc, sizeof(float) * (size_t)C);
                } else if (op == T_AXPYST_S) {
                    const float* b = w.read(slot, chunk);
                    for (int e = 0; e < C; e++) { float t = acc[e] * im[0]; t = t + im[1]; t = t * im[2]; acc[e] = t + b[e]; }
                    const float* c2 = w.read(ex[3], chunk);          // read BEFORE the reload: the two operands never share a slot
                    if (ex[3] == slot) bad("T_AXPYST with both operands in one slot");
                    for (int e = 0; e < C; e++) { const float t = c2[e] * im[3]; acc[e] = acc[e] + t; }
                    if ((int)ey[4] >= P.n_ptrs) bad("pointer index %d out of range", (int)ey[4]);
                    float* dst = P.ptrs[ey[4]];
                    stored.insert(dst);
                    for (int e = 0; e < C && base + e < P.n; e++) dst[base + e] = acc[e];
                    if (ey[5] != 0xffffffffu) {
                        if ((int)slot >= P.n_ring) bad("T_AXPYST re-arms a slot outside the ring");
                        w.load(slot, ey[5], chunk);
                    }
                } else {
                    if (op == T_ADDAFFDISC_W || op == T_ADDAFFDISC_WL) w.wait(slot);
                    const float* b = w.read(slot, chunk);
                    for (int e = 0; e < C; e++) {
                        float t = b[e] + im[0]; t = t * im[1]; const float num = acc[e] + t;
                        float d = b[e] * im[2]; d = d + 1.0f;
                        acc[e] = num / d;
                    }
                    if (op == T_ADDAFFDISC_SL || op == T_ADDAFFDISC_WL) w.load(slot, ey[3], chunk);
                }
                pc += ext;
                continue;
            }
            if (op == T_ADDMUL_II) {
                if (pc + 1 >= P.n_instr) bad("two-word instruction at the end of the tape");
                if (P.instr[pc + 1].x != T_END) bad("the word behind two-word opcode %d is not its extension word", (int)op);
                float imm2; std::memcpy(&imm2, &P.instr[pc + 1].y, 4);
                for (int e = 0; e < C; e++) { const float t = acc[e] + imm; acc[e] = t * imm2; }
                pc++;
                continue;
            }
            if (op >= T_BIN0) {
                if (op >= T_NUM_OPS) bad("opcode %d out of range", (int)op);
                const uint32_t kk = (op - T_BIN0) / 3u, fl = (op - T_BIN0) % 3u;
                const float* b = nullptr;
                if (fl == 2u) { w.wait(slot); b = w.read(slot, chunk); }
                else if (fl == 1u) b = w.read(slot, chunk);
                else if (kk >= 10u) bad("compound op %d has no immediate-operand form", (int)kk);
                for (int e = 0; e < C; e++) {
                    const float x = b ? b[e] : imm;
                    float& a = acc[e];
                    switch (kk) {
                    case 0: a = x; break;
                    case 1: a = a + x; break;
                    case 2: a = a - x; break;
                    case 3: a = x - a; break;
                    case 4: a = a * x; break;
                    case 5: a = a / x; break;
                    case 6: a = x / a; break;
                    case 7: a = jminf(a, x); break;
                    case 8: a = jmaxf(a, x); break;
                    case 9: a = pred[e] ? a : x; break;
                    case 10: { const float t = x * imm; a = a + t; break; }
                    case 11: { float t = x * imm; t = t + 1.0f; a = a * t; break; }
                    case 12: { float t = x * imm; t = t + 1.0f; a = a / t; break; }
                    default: bad("binary op %d unknown", (int)kk);
                    }
                }
                continue;
            }
            switch (op) {
            case T_LOAD: w.load(slot, in.y, chunk); break;
            case T_LOADN: if (next < n_chunks) w.load(slot, in.y, next); break;
            case T_WAIT: w.wait(slot); break;
            case T_STG: case T_STGS: {
                if ((int)in.y >= P.n_ptrs) bad("pointer index %d out of range", (int)in.y);
                const float* v = acc;
                if (op == T_STGS) { if ((int)slot < P.n_ring) bad("T_STGS from a ring slot"); v = w.read(slot, chunk); }
                float* dst = P.ptrs[in.y];
                stored.insert(dst);
                for (int e = 0; e < C && base + e < P.n; e++) dst[base + e] = v[e];
                break;
            }
            case T_STR:
                if ((int)slot < P.n_ring || (int)slot >= P.n_slots) bad("T_STR into slot %d outside the register file", (int)slot);
                std::memcpy(&w.slots[(size_t)slot * C], acc, sizeof(float) * (size_t)C);
                w.reg_written[slot] = 1;
                break;
            case T_SETP: for (int e = 0; e < C; e++) pred[e] = acc[e] >= 0.0f; break;
            case T_SQR: for (int e = 0; e < C; e++) acc[e] = acc[e] * acc[e]; break;
            case T_SQRT: for (int e = 0; e < C; e++) acc[e] = std::sqrt(acc[e]); break;
            case T_EXP: for (int e = 0; e < C; e++) acc[e] = (float)std::exp((double)acc[e]); break;
            case T_LOG: for (int e = 0; e < C; e++) acc[e] = (float)std::log((double)acc[e]); break;
            case T_SIN: for (int e = 0; e < C; e++) acc[e] = (float)std::sin((double)acc[e]); break;
            case T_COS: for (int e = 0; e < C; e++) acc[e] = (float)std::cos((double)acc[e]); break;
            case T_ABS: for (int e = 0; e < C; e++) acc[e] = std::fabs(acc[e]); break;
            case T_INV: for (int e = 0; e < C; e++) acc[e] = 1.0f / acc[e]; break;
            case T_ISNAN: for (int e = 0; e < C; e++) acc[e] = (acc[e] != acc[e]) ? 1.0f : 0.0f; break;
            case T_POW: for (int e = 0; e < C; e++) acc[e] = jpow(acc[e], imm); break;
            case T_MULADD_II: {
                if (pc + 1 >= P.n_instr) bad("two-word instruction at the end of the tape");
                if (P.instr[pc + 1].x != T_END) bad("the word behind two-word opcode %d is not its extension word", (int)op);
                float imm2; std::memcpy(&imm2, &P.instr[pc + 1].y, 4);
                for (int e = 0; e < C; e++) { const float t = acc[e] * imm; acc[e] = t + imm2; }
                pc++;                                   // the extension word is not an instruction
                break;
            }
            case T_ACCUM_S: {
                if ((int)slot < P.n_ring) bad("T_ACCUM_S on a ring slot");
                const float* b = w.read(slot, chunk);
                for (int e = 0; e < C; e++) acc[e] = acc[e] + b[e];
                std::memcpy(&w.slots[(size_t)slot * C], acc, sizeof(float) * (size_t)C);
                break;
            }
            default: bad("opcode %d unknown", (int)op);
            }
        }
        // only T_LOADN'ed copies (for the chunk that uses this slot set next) may still be in flight
        for (int s = 0; s < P.n_ring; s++) {
            if (w.issued[s] != w.waited[s]) {
                if (next >= n_chunks) bad("ring slot %d has a copy in flight when the warp is done with its slot set", s);
                if (w.slot_chunk[s] != next) bad("ring slot %d carries a copy of the CURRENT chunk across the chunk boundary", s);
            }
        }
        if (P.reduce_mode != RM_NONE) {
            for (int e = 0; e < C && base + e < P.n; e++) {
                const double x = (double)acc[e];
                switch (P.reduce_mode) {
                case RM_SUM: part.s += x; break;
                case RM_MOMENTS: values.push_back(x); break;
                case RM_MIN: part.mn = part.c == 0 ? x : (double)jminf((float)part.mn, (float)x); break;      // NaN sticks, -0 < +0
                case RM_MAX: part.mx = part.c == 0 ? x : (double)jmaxf((float)part.mx, (float)x); break;
                case RM_DOT: if (!endb) bad("RM_DOT without a slot operand on T_END"); part.s += (double)endb[e] * x; break;
                case RM_WSQ: { if (!endb) bad("RM_WSQ without a slot operand on T_END"); const double d = (double)endb[e] - P.reduce_param; part.s += d * d * x; break; }
                default: bad("reduce mode %d unknown", P.reduce_mode);
                }
                part.c += 1.0;
            }
        }
    }
}

void dump_tape(const TapeParams& P, int grid) {
    static const char* names[] = {"END", "LOAD", "WAIT", "STG", "STGS", "STR", "SETP", "SQR", "SQRT", "EXP", "LOG", "SIN", "COS", "ABS", "INV",
                                  "ISNAN", "POW", "MULADD_II", "LOADN", "ACCUM_S", "?"};
    static const char* bins[] = {"MOV", "ADD", "SUB", "BUS", "MUL", "DIV", "VID", "MIN", "MAX", "SEL", "ADDPROD", "ACCRUE", "DISCOUNT"};
    std::fprintf(stderr, "[tape] n=%lld elems=%d grid=%d instr=%d prologue=%d ptrs=%d ring=%d slots=%d reduce=%d\n", P.n, P.elems, grid, P.n_instr, P.n_prologue,
                 P.n_ptrs, P.n_ring, P.n_slots, P.reduce_mode);
    const int SHIFT = tape_slot_shift(P.elems);
    for (int i = 0; i < P.n_instr; i++) {
        const uint32_t op = P.instr[i].x & ((1u << SHIFT) - 1u), slot = P.instr[i].x >> SHIFT;
        float imm; std::memcpy(&imm, &P.instr[i].y, 4);
        if (op == T_MULADDMUL) std::fprintf(stderr, "  %4d MULADDMUL %g\n", i, imm);
        else if (op == T_RATIO) std::fprintf(stderr, "  %4d RATIO %g\n", i, imm);
        else if (op == T_ADDAFFDISC_SL || op == T_ADDAFFDISC_WL) std::fprintf(stderr, "  %4d ADDAFFDISC_%cL s%u %g\n", i, op == T_ADDAFFDISC_SL ? 'S' : 'W', slot, imm);
        else if (op == T_RATIOACC_S || op == T_RATIOACC_W) std::fprintf(stderr, "  %4d RATIOACC_%c s%u %g\n", i, op == T_RATIOACC_S ? 'S' : 'W', slot, imm);
        else if (op == T_AXPYST_S) std::fprintf(stderr, "  %4d AXPYST_S s%u %g\n", i, slot, imm);
        else if (op == T_RATIOACC_A) std::fprintf(stderr, "  %4d RATIOACC_A s%u %g\n", i, slot, imm);
        else if (op == T_ADDAFFDISC_S || op == T_ADDAFFDISC_W) std::fprintf(stderr, "  %4d ADDAFFDISC_%c s%u %g\n", i, op == T_ADDAFFDISC_S ? 'S' : 'W', slot, imm);
        else if (op == T_ADDMUL_II) std::fprintf(stderr, "  %4d ADDMUL_II %g\n", i, imm);
        else if (op == T_ADDAFF_S || op == T_ADDAFF_W) std::fprintf(stderr, "  %4d ADDAFF_%c s%u %g\n", i, op == T_ADDAFF_S ? 'S' : 'W', slot, imm);
        else if (op >= T_BIN0) {
            const uint32_t k = (op - T_BIN0) / 3u, fl = (op - T_BIN0) % 3u;
            if (fl == 0) std::fprintf(stderr, "  %4d %s_I %g\n", i, bins[k], imm);
            else std::fprintf(stderr, "  %4d %s_%c s%u%s (imm %g)\n", i, bins[k], fl == 1 ? 'S' : 'W', slot, (int)slot < P.n_ring ? "" : "r", imm);
        } else std::fprintf(stderr, "  %4d %s s%u y=%u\n", i, names[op < 20 ? op : 20], slot, P.instr[i].y);
    }
}

}  // namespace

cudaError_t launch_tape(const TapeParams& P, int grid, int n_warps, cudaStream_t, cudaStream_t) {
    if (std::getenv("FMC_EMU_DUMP")) dump_tape(P, grid);
    static const bool noexec = std::getenv("FMC_EMU_NOEXEC") != nullptr;    // host-side profiling: launches cost nothing
    if (noexec) {
        if (P.reduce_mode != RM_NONE) {
            P.result[0] = (double)P.n; P.result[1] = 0.0; P.result[2] = 0.0;
            if (P.host_result) { P.host_result[0] = (double)P.n; P.host_result[1] = 0.0; P.host_result[2] = 0.0; P.host_result[3] = P.ticket; }
        }
        return cudaSuccess;
    }
    try {
        if (!tape_valid_elems(P.elems)) bad("bad chunk geometry: %d elements per lane", P.elems);
        if (P.n_ring < 0 || P.n_ring > TAPE_MAX_RING || P.n_slots < P.n_ring) bad("bad slot counts: ring %d slots %d", P.n_ring, P.n_slots);
        if (P.n_instr < 1 || P.n_instr > TAPE_MAX_INSTR + 1) bad("bad instruction count %d", P.n_instr);
        if (P.n_prologue < 0 || P.n_prologue + 1 >= P.n_instr) bad("bad prologue length %d", P.n_prologue);
        if (P.n_sets < 1 || P.n_sets > 4) bad("bad slot-set count %d", P.n_sets);
        if (grid < 1) bad("empty grid");
        if (tape_smem_bytes(P.n_ptrs, P.n_instr, P.n_slots, P.n_sets, n_warps, P.elems) > 232448 - 1024) bad("shared memory of one CTA exceeds the device limit");
        std::set<const float*> stored;
        Partial part;
        std::vector<double> values;
        if (n_warps < 1 || n_warps > TAPE_MAX_WARPS) bad("bad warp count %d", n_warps);
        const long long stride = (long long)grid * n_warps;
        for (long long w = 0; w < stride; w++) run_warp(P, w, stride, stored, part, values);
        if (P.reduce_mode != RM_NONE) {
            double v = part.s, m2 = 0.0;
            if (P.reduce_mode == RM_MIN) v = part.mn;
            else if (P.reduce_mode == RM_MAX) v = part.mx;
            else if (P.reduce_mode == RM_MOMENTS) {
                double s = 0.0; for (double x : values) s += x;
                const double mean = values.empty() ? 0.0 : s / (double)values.size();
                for (double x : values) m2 += (x - mean) * (x - mean);
                v = mean; part.c = (double)values.size();
            }
            P.result[0] = part.c; P.result[1] = v; P.result[2] = m2;
            // (sums of an unsharded vector: the kernel publishes {host[2] = value, host[3] = ticket}, reduce_common.cuh)
            const bool sum_pair = P.xchg.nranks <= 1 && (P.reduce_mode == RM_SUM || P.reduce_mode == RM_DOT || P.reduce_mode == RM_WSQ);
            if (P.host_result) { P.host_result[0] = part.c; P.host_result[1] = v; P.host_result[2] = sum_pair ? v : m2; P.host_result[3] = P.ticket; }
        }
        return cudaSuccess;
    } catch (const Check& c) {
        g_emu_error = c.msg;
        std::fprintf(stderr, "[tape emulator] ILLEGAL TAPE: %s\n", c.msg.c_str());
        return cudaErrorLaunchFailure;
    }
}
void tape_kernel_teardown() {}
cudaError_t launch_cast_f64_f32(const double* src, float* dst, long long n, int, cudaStream_t) {
    for (long long i = 0; i < n; i++) dst[i] = (float)src[i];
    return cudaSuccess;
}
cudaError_t tape_kernel_setup(size_t* m) { if (m) *m = 232448 - 1024; return cudaSuccess; }
size_t tape_smem_bytes(int n_ptrs, int n_instr, int n_slots, int n_sets, int n_warps, int elems) {
    size_t s = (size_t)n_warps * (size_t)n_sets * TAPE_MAX_RING * 8;
    s += ((size_t)n_ptrs * 8 + 15) & ~(size_t)15;
    s = (s + ((size_t)n_instr + 2) * 8 + 127) & ~(size_t)127;
    return s + (size_t)n_warps * (size_t)n_sets * (size_t)n_slots * (size_t)tape_slot_bytes(elems);
}
int tape_max_blocks_per_sm(size_t smem_bytes, int, int n_warps, int elems) {
    const int by_smem = (int)((233472 - 1024) / (smem_bytes + 1024));
    const int regs = elems == 16 ? 128 : elems == 8 ? 72 : 40;            // __maxnreg__ of the three geometries (tape_interp.cuh)
    const int by_regs = 65536 / (regs * 32 * n_warps);
    return std::max(1, std::min(by_smem, by_regs));
}

// the other kernels are not ISA-driven; the emulator gives the regression its plain meaning and declines the rest
cudaError_t launch_regression(const RegressionParams& P, int, cudaStream_t) {
    const int k = P.k;
    std::vector<double> acc((size_t)(k * (k + 1) / 2 + k), 0.0);
    for (long long p = 0; p < P.n; p++) {
        float b[REG_MAX_K];
        for (int i = 0; i < k; i++) b[i] = P.basis[i] ? P.basis[i][p] : P.scalars[i];
        int t = 0;
        for (int i = 0; i < k; i++) for (int j = i; j < k; j++, t++) { const float x = b[i] * b[j]; acc[t] += P.float_products ? (double)x : (double)b[i] * (double)b[j]; }
        for (int i = 0; i < k; i++, t++) { const float x = P.y[p] * b[i]; acc[t] += P.float_products ? (double)x : (double)P.y[p] * (double)b[i]; }
    }
    for (size_t t = 0; t < acc.size(); t++) P.result[t] = acc[t];
    return cudaSuccess;
}
int regression_max_blocks_per_sm(int) { return 1; }
static uint32_t emu_sort_key(float f) {
    if (f != f) return 0xffffffffu;
    uint32_t u; std::memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
cudaError_t launch_select_hist(const float* x, long long n, uint32_t prefix, uint32_t mask, int shift, double* hist, int, cudaStream_t) {
    for (long long i = 0; i < n; i++) { const uint32_t k = emu_sort_key(x[i]); if (((k ^ prefix) & mask) == 0u) hist[(k >> shift) & 255u] += 1.0; }
    return cudaSuccess;
}
cudaError_t launch_range_stats(const float* x, long long n, uint32_t lo, uint32_t hi, double* out, int, cudaStream_t) {
    for (long long i = 0; i < n; i++) {
        const uint32_t k = emu_sort_key(x[i]);
        if (k < lo) out[0] += 1.0; else if (k == lo) out[1] += 1.0; else if (k < hi) { out[2] += 1.0; out[3] += (double)x[i]; } else if (k == hi) out[4] += 1.0;
    }
    return cudaSuccess;
}
cudaError_t launch_histogram(const float* x, long long n, const double* pts, int m, double* counts, int, cudaStream_t) {
    for (long long i = 0; i < n; i++) { int k = 0; while (k < m && !((double)x[i] <= pts[k])) k++; counts[k] += 1.0; }
    return cudaSuccess;
}
int regression_tile_elems() { return 2048; }
cudaError_t launch_reduce(const ReduceParams& P, int, cudaStream_t) {
    double c = 0, s = 0, m2 = 0, mn = 0, mx = 0;
    std::vector<double> vals;
    for (long long i = 0; i < P.n; i++) {
        const double x = (double)P.x[i], w = P.w ? (double)P.w[i] : 0.0;
        switch (P.mode) {
        case RM_SUM: s += x; break;
        case RM_MOMENTS: vals.push_back(x); break;
        case RM_MIN: mn = c == 0 ? x : (double)jminf((float)mn, (float)x); break;
        case RM_MAX: mx = c == 0 ? x : (double)jmaxf((float)mx, (float)x); break;
        case RM_DOT: s += x * w; break;
        case RM_WSQ: s += (x - P.param) * (x - P.param) * w; break;
        default: return cudaErrorInvalidValue;
        }
        c += 1.0;
    }
    double v = s;
    if (P.mode == RM_MIN) v = mn; else if (P.mode == RM_MAX) v = mx;
    else if (P.mode == RM_MOMENTS) {
        double t = 0; for (double x : vals) t += x;
        v = vals.empty() ? 0.0 : t / (double)vals.size();
        for (double x : vals) m2 += (x - v) * (x - v);
    }
    P.result[0] = c; P.result[1] = v; P.result[2] = m2;
    const bool sum_pair = P.xchg.nranks <= 1 && (P.mode == RM_SUM || P.mode == RM_DOT || P.mode == RM_WSQ);
    if (P.host_result) { P.host_result[0] = c; P.host_result[1] = v; P.host_result[2] = sum_pair ? v : m2; P.host_result[3] = P.ticket; }
    return cudaSuccess;
}
cudaError_t launch_batch_sum(const BatchSumParams& P, cudaStream_t) {
    if (std::getenv("FMC_EMU_NOEXEC")) { P.host_out[BATCH_MAX] = P.ticket; return cudaSuccess; }
    for (int j = 0; j < P.k; j++) {
        double s = 0;
        for (long long i = 0; i < P.n; i++) s += (double)P.x[j][i];
        P.host_out[j] = s;
    }
    P.host_out[BATCH_MAX] = P.ticket;
    return cudaSuccess;
}
int reduce_tile_elems() { return 4096; }
// FMC_EMU_FAKE_BROWNIAN=1 (tape-shape studies of the workload drivers only): increments from a throw-away generator,
// NOT the MT19937 stream — the Brownian parity tests are never run against the emulator.
static bool fake_brownian() { return std::getenv("FMC_EMU_FAKE_BROWNIAN") != nullptr; }
int brownian_max_blocks_per_sm(int, int, int) { return 4; }
cudaError_t launch_brownian(const BrownianParams& P, cudaStream_t) {
    if (!fake_brownian()) return cudaErrorNotSupported;
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int v = 0; v < P.T * P.F; v++)
        for (long long p = 0; p < P.np; p++) {
            double u = 0.0;
            for (int k = 0; k < 12; k++) { s = s * 6364136223846793005ull + 1442695040888963407ull; u += (double)(s >> 11) * (1.0 / 9007199254740992.0); }
            P.out[v][p] = (float)((u - 6.0) * P.sqrt_dt[v / P.F]);
        }
    return cudaSuccess;
}
cudaError_t launch_mt_jump(const uint32_t*, const uint32_t*, int, const long long*, uint32_t*, int, cudaStream_t) { return fake_brownian() ? cudaSuccess : cudaErrorNotSupported; }
cudaError_t launch_mt_raw(const uint32_t*, const long long*, int, long long, unsigned long long, long long, uint32_t*, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace fmc

extern "C" const char* fmc_emu_last_tape_error(void) { return fmc::g_emu_error.c_str(); }
