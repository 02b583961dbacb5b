// codegen.cpp — turns the pending-op graph into interpreter tapes.
//
// A flush takes a set of target nodes, collects their cone of pending (lazy) nodes, decides which nodes must be
// written to HBM (still referenced by the caller or by pending nodes outside the cone) and which live only in
// registers, allocates the interpreter's registers (accumulator machine, see tape_isa.h) and launches as few
// tape kernels as the instruction / pointer-table limits allow. A value that does not fit in the register
// file is spilled to a pooled buffer with an ordinary store and re-read as a leaf (a thread reads back its own
// write), so register pressure never forces a kernel boundary.
//
// Invariant: every node is computed exactly once, except the private chain of a reduction target
// ("ephemeral" nodes), which is evaluated in registers for the reduction and stays pending.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>

#include "runtime.h"

namespace fmc {

namespace {

// FMC_HOST_PROFILE=1: wall time per code-generation phase, printed at exit (development aid)
struct PhaseClock {
    double us[8] = {0}; const char* name[8] = {"collect", "classify", "emit", "peephole", "schedule", "launch-prep", "bookkeeping", "other"};
    bool on = std::getenv("FMC_HOST_PROFILE") != nullptr;
    ~PhaseClock() { if (on) for (int i = 0; i < 8; i++) std::fprintf(stderr, "[fmc host] %-12s %10.1f us\n", name[i], us[i]); }
};
PhaseClock g_phase;
struct PhaseTimer {
    int id; std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(int i) : id(i), t0(std::chrono::steady_clock::now()) {}
    ~PhaseTimer() { if (g_phase.on) g_phase.us[id] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
};

constexpr int32_t NO_USE = 0x7fffffff;

struct Info {
    int32_t node;
    int32_t uses = 0;       // operand slots that read this value inside the flush (+1 for the reduction epilogue)
    int32_t eph_uses = 0;
    int32_t ubeg = 0, ucur = 0, uend = 0;   // window into Gen::use_list (positions of the consuming nodes, ascending)
    int16_t slot = -1;      // pointer-table slot in the current kernel
    int8_t reg = -1;
    bool lazy = false;      // cone node (to be computed) vs. materialised leaf
    bool store = false, eph = false, computed = false;
    bool written_here = false;   // its HBM buffer is written by the kernel being built: no TMA read-back before the next kernel
    float* buf = nullptr;   // device buffer (existing for leaves, new for stored / spilled nodes)
};

// ---- abstract accumulator-machine code (before ring scheduling) ----
enum AKind : uint8_t { K_NONE = 0, K_IMM, K_REG, K_LEAF, K_RELOAD /* extension word: pointer-table slot of the leaf a fused form re-arms its ring slot with */ };
enum BinOp : uint16_t { B_MOV = 0, B_ADD, B_SUB, B_BUS, B_MUL, B_DIV, B_VID, B_MIN, B_MAX, B_SEL, B_ADDPROD, B_ACCRUE, B_DISCOUNT };
constexpr uint16_t A_BIN = 0x100;   // AIns::op = A_BIN | BinOp for binary instructions, a plain TapeOp otherwise
constexpr uint16_t A_EXT = 0x200;   // extension word of the two-word instruction before it: y = second immediate
struct AIns {
    uint16_t op;
    uint8_t kind;
    int32_t arg;      // K_REG: register-file slot; K_LEAF: local id of the leaf
    uint32_t y;       // immediate bits / pointer-table slot
    // where the immediates came from (tape cache: a replay patches them): 4 * cone position + operand index, -1 = not an
    // immediate; neg: the word holds the negated value (SUB_I a == ADD_I -a). src2 belongs to the second immediate of the
    // two-word forms, which travels in `arg`.
    int32_t src = -1, src2 = -1;
    uint8_t neg = 0;
};
AIns mk(uint16_t op, uint8_t kind, int32_t arg, uint32_t y, int32_t src = -1, uint8_t neg = 0, int32_t src2 = -1) {
    AIns a; a.op = op; a.kind = kind; a.arg = arg; a.y = y; a.src = src; a.neg = neg; a.src2 = src2; return a;
}

// ---- tape cache: the launches of one cone, keyed by the cone's structure (see Runtime::run_cone) ----
struct ImmPatch { int32_t word; int32_t src; uint8_t neg; };
struct KernelPlan {
    int n_instr = 0, n_prologue = 0, n_ring = 0, n_slots = 0, n_sets = 1, reduce_mode = RM_NONE, grid = 1, n_warps = TAPE_WARPS;
    int elems = TAPE_E_MAX;             // chunk geometry (elements per lane) the words are encoded for
    int n_leaf_slots = 0, n_result_stores = 0;
    std::vector<TapeInstr> words;       // n_instr + 2
    std::vector<int32_t> ptr_local;     // pointer table: the local whose buffer goes into each entry
    std::vector<ImmPatch> patches;
};
struct ConePlan {
    std::vector<uint32_t> key;
    std::vector<KernelPlan> kernels;
    std::vector<uint8_t> node_flags;    // per cone node: bit 0 gets a buffer (stored or spilled), bit 1 ephemeral
};
struct TapeCache {
    std::unordered_map<uint64_t, std::vector<std::unique_ptr<ConePlan>>> map;
    size_t entries = 0, words = 0;
    uint64_t hits = 0, misses = 0;
    void clear() { map.clear(); entries = 0; words = 0; }
};
TapeCache g_cache;

uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

uint16_t rev_op(uint16_t op) {
    switch (op) {
    case B_SUB: return B_BUS;
    case B_BUS: return B_SUB;
    case B_DIV: return B_VID;
    case B_VID: return B_DIV;
    default: return op;   // ADD MUL MIN MAX commute
    }
}

struct OccCache { std::map<std::pair<size_t, int>, int> blocks; };

struct Gen {
    Runtime& rt;
    int64_t n;
    std::vector<Info> info;
    std::vector<int32_t> use_list;
    std::vector<AIns> A;                // abstract code of the kernel being built
    std::vector<float*> ptrs;
    std::vector<int32_t> slotted;       // locals that own a pointer-table slot in the current kernel
    std::vector<int32_t> written;       // locals whose buffer the current kernel writes
    int32_t reg_owner[TAPE_REGS];
    int32_t acc_owner = -1;
    int32_t pos = 0;                    // cone position of the node being emitted
    int regs_used = 0, n_leaf_refs = 0;
    int n_leaf_slots = 0, n_result_stores = 0;   // algorithmic traffic of the current kernel (spills / re-reads excluded)
    TapeParams* params;                 // reused launch parameter block

    Gen(Runtime& r, int64_t n_) : rt(r), n(n_) {
        for (int j = 0; j < TAPE_REGS; j++) reg_owner[j] = -1;
        static thread_local TapeParams tp;
        params = &tp;
    }

    bool has_buf(int32_t L) const { return info[L].buf != nullptr; }
    bool reloadable(int32_t L) const { return info[L].buf != nullptr && !info[L].written_here; }
    int32_t rem(int32_t L) const { return info[L].uend - info[L].ucur; }
    int32_t next_use(int32_t L) const { const Info& f = info[L]; return f.ucur < f.uend ? use_list[f.ucur] : NO_USE; }
    // next use that is not the instruction being built (which consumes `now` operand slots)
    int32_t next_use_after(int32_t L, int now) const { const Info& f = info[L]; return f.ucur + now < f.uend ? use_list[f.ucur + now] : NO_USE; }
    bool available(int32_t L) const { return acc_owner == L || info[L].reg >= 0 || reloadable(L); }

    void begin_kernel() {
        A.clear(); ptrs.clear();
        for (int32_t L : slotted) info[L].slot = -1;
        slotted.clear();
        for (int32_t L : written) info[L].written_here = false;
        written.clear();
        for (int j = 0; j < TAPE_REGS; j++) {
            if (reg_owner[j] >= 0) info[reg_owner[j]].reg = -1;
            reg_owner[j] = -1;
        }
        acc_owner = -1;
        regs_used = 0; n_leaf_refs = 0;
        n_leaf_slots = 0; n_result_stores = 0;
    }

    int slot_for(int32_t L) {
        Info& f = info[L];
        if (f.slot < 0) {
            f.slot = (int16_t)ptrs.size();
            ptrs.push_back(f.buf);
            slotted.push_back(L);
            // (algorithmic traffic: a value an earlier window of the same flush stored is a re-read, not an input)
            if (!f.lazy && !(rt.windowing && rt.window_stored.count(f.buf))) n_leaf_slots++;
        }
        return f.slot;
    }

    void ensure_buffer(int32_t L) {
        Info& f = info[L];
        if (!f.buf) f.buf = (float*)rt.pool.alloc(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
    }

    void emit(uint16_t op, uint8_t kind = K_NONE, int32_t arg = 0, uint32_t y = 0, int32_t src = -1) { A.push_back(mk(op, kind, arg, y, src)); }
    void emit_src(uint16_t binop, int32_t L) {      // binary instruction whose operand is the value L (register file or leaf)
        const Info& f = info[L];
        if (f.reg >= 0) emit(A_BIN | binop, K_REG, f.reg);
        else if (reloadable(L)) { slot_for(L); emit(A_BIN | binop, K_LEAF, L); n_leaf_refs++; }
        else fail(FMC_ERR_UNSUPPORTED, "internal: operand %d has no location (at %d: lazy %d computed %d store %d buffer %d written here %d, uses left %d)",
                  (int)L, (int)pos, (int)f.lazy, (int)f.computed, (int)f.store, (int)(f.buf != nullptr), (int)f.written_here, (int)rem(L));
    }

    // write local L (currently in acc or in a register) to its HBM buffer
    void store_value(int32_t L) {
        ensure_buffer(L);
        const int slot = slot_for(L);
        if (acc_owner == L) emit(T_STG, K_NONE, 0, (uint32_t)slot);
        else emit(T_STGS, K_REG, info[L].reg, (uint32_t)slot);
        if (!info[L].written_here) { info[L].written_here = true; written.push_back(L); }
    }

    void free_reg_of(int32_t L) {
        Info& f = info[L];
        if (f.reg >= 0) { reg_owner[f.reg] = -1; f.reg = -1; }
    }

    int32_t pins[3] = {-1, -1, -1};      // operands of the instruction being built: never evicted
    void set_pins(int32_t a, int32_t b = -1, int32_t c = -1) { pins[0] = a; pins[1] = b; pins[2] = c; }

    // Find a register-file slot. When all are taken, the value whose next use is farthest away leaves: one that already
    // has an HBM copy is simply dropped, any other is stored first. Either way it cannot be read back through the TMA
    // ring inside this kernel (its buffer is written here), so its next use ends the kernel (see emit loop).
    int alloc_reg() {
        const int n_regs = std::max(4, std::min(rt.opt.max_regs, TAPE_REGS));
        for (int j = 0; j < n_regs; j++) if (reg_owner[j] < 0) { regs_used = std::max(regs_used, j + 1); return j; }
        int victim = -1, victim_buf = -1;
        int32_t far = -1, far_buf = -1;
        for (int j = 0; j < n_regs; j++) {
            const int32_t L = reg_owner[j];
            if (L == pins[0] || L == pins[1] || L == pins[2]) continue;
            const int32_t nu = next_use(L);
            if (nu > far) { far = nu; victim = j; }
            if (has_buf(L) && nu > far_buf) { far_buf = nu; victim_buf = j; }
        }
        if (victim_buf >= 0 && far_buf - pos > 32) victim = victim_buf;
        if (victim < 0) fail(FMC_ERR_UNSUPPORTED, "internal: register allocation failed");
        const int32_t L = reg_owner[victim];
        if (!has_buf(L)) store_value(L);        // spill
        free_reg_of(L);
        return victim;
    }

    void put_acc_in_reg() {
        const int32_t L = acc_owner;
        const int j = alloc_reg();
        emit(T_STR, K_REG, j);
        reg_owner[j] = L; info[L].reg = (int8_t)j;
    }
    bool free_reg_exists() const {
        const int n_regs = std::max(4, std::min(rt.opt.max_regs, TAPE_REGS));
        for (int j = 0; j < n_regs; j++) if (reg_owner[j] < 0) return true;
        return false;
    }

    // the accumulator is about to be overwritten: keep its value if somebody still needs it
    void save_acc() {
        const int32_t L = acc_owner;
        if (L < 0) return;
        if (rem(L) > 0 && info[L].reg < 0) {
            if (!has_buf(L)) put_acc_in_reg();
            // an operand of the instruction being built that cannot come back through the ring (its buffer is written by
            // this kernel) must stay on chip, whatever it costs
            else if ((L == pins[0] || L == pins[1] || L == pins[2]) && !reloadable(L)) put_acc_in_reg();
            // the value is already in HBM: a register copy pays only for a use that is near
            else if (next_use(L) - pos < 192 && free_reg_exists()) put_acc_in_reg();
        }
    }

    // make acc hold L; `uses_now` operand slots of L are consumed by the instruction being built
    void take_acc(int32_t L, int uses_now) {
        if (acc_owner != L) {
            save_acc();
            emit_src(B_MOV, L);
            acc_owner = L;
        }
        const Info& f = info[L];
        if (rem(L) > uses_now && f.reg < 0) {
            if (!has_buf(L)) put_acc_in_reg();
            else if (!reloadable(L) && next_use_after(L, uses_now) - pos < 192 && free_reg_exists()) put_acc_in_reg();
        }
    }

    void consume(int32_t L) {
        Info& f = info[L];
        f.ucur++;
        if (f.ucur >= f.uend) free_reg_of(L);
    }

    void emit_binary(uint16_t op, int32_t A_, float immA, int32_t B_, float immB) {
        // choose the operand that sits in (or goes to) the accumulator
        int32_t X, O; float immO; uint16_t xop; int kO;
        if (A_ >= 0 && acc_owner == A_) { X = A_; O = B_; immO = immB; xop = op; kO = 1; }
        else if (B_ >= 0 && acc_owner == B_) { X = B_; O = A_; immO = immA; xop = rev_op(op); kO = 0; }
        else if (A_ >= 0) { X = A_; O = B_; immO = immB; xop = op; kO = 1; }
        else { X = B_; O = A_; immO = immA; xop = rev_op(op); kO = 0; }
        set_pins(X, O);
        if (O == X) {
            // both operands are the same value
            take_acc(X, 2);
            if (op == B_MUL) emit(T_SQR);
            else {
                if (info[X].reg < 0) put_acc_in_reg();
                emit(A_BIN | xop, K_REG, info[X].reg);
            }
            consume(X); consume(X);
            return;
        }
        take_acc(X, 1);
        if (O >= 0) emit_src(xop, O);
        else emit(A_BIN | xop, K_IMM, 0, f2u(immO), 4 * pos + kO);
        consume(X);
        if (O >= 0) consume(O);
    }

    void emit_node(int32_t L) {
        pos = L;
        const Node& nd = rt.nodes[info[L].node];
        auto loc = [&](int k) -> int32_t { return nd.in[k] >= 0 ? rt.nodes[nd.in[k]].local : -1; };
        switch (nd.op) {
        case N_ADD: emit_binary(B_ADD, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_SUB: emit_binary(B_SUB, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MUL: emit_binary(B_MUL, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_DIV: emit_binary(B_DIV, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MIN: emit_binary(B_MIN, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MAX: emit_binary(B_MAX, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_SQRT: case N_EXP: case N_LOG: case N_SIN: case N_COS: case N_ABS: case N_INV: case N_ISNAN: case N_POW: {
            static const uint16_t map[] = {0, 0, 0, 0, 0, 0, 0, T_SQRT, T_EXP, T_LOG, T_SIN, T_COS, T_ABS, T_INV, T_ISNAN, T_POW};
            const int32_t A_ = loc(0);
            set_pins(A_);
            take_acc(A_, 1);
            emit(map[nd.op], K_NONE, 0, f2u(nd.imm[1]), 4 * L + 1);
            consume(A_);
            break;
        }
        case N_CHOOSE: {
            const int32_t X = loc(0), A_ = loc(1), B_ = loc(2);
            set_pins(X, A_, B_);
            take_acc(X, 1);
            emit(T_SETP);
            consume(X);
            // acc := A
            if (A_ >= 0) {
                if (acc_owner != A_) {
                    save_acc();
                    emit_src(B_MOV, A_);
                    acc_owner = A_;
                }
                const int now = (B_ == A_) ? 2 : 1;
                if (rem(A_) > now && info[A_].reg < 0 && !reloadable(A_)) put_acc_in_reg();
                else if (B_ == A_ && info[A_].reg < 0 && !reloadable(A_)) put_acc_in_reg();
            } else {
                save_acc();
                emit(A_BIN | B_MOV, K_IMM, 0, f2u(nd.imm[1]), 4 * L + 1);
                acc_owner = -1;
            }
            if (B_ >= 0) {
                if (B_ == A_) { /* choose(x, a, a) == a */ }
                else emit_src(B_SEL, B_);
            } else {
                emit(A_BIN | B_SEL, K_IMM, 0, f2u(nd.imm[2]), 4 * L + 2);
            }
            if (A_ >= 0) consume(A_);
            if (B_ >= 0) consume(B_);
            break;
        }
        case N_CONST:
            set_pins(-1);
            save_acc();
            emit(A_BIN | B_MOV, K_IMM, 0, f2u(nd.imm[0]), 4 * L + 0);
            break;
        default: fail(FMC_ERR_UNSUPPORTED, "internal: cannot emit node op %d", (int)nd.op);
        }
        acc_owner = L;
        info[L].computed = true;
        if (info[L].store) { store_value(L); n_result_stores++; }
    }

    // every operand of node L can be fetched inside the kernel being built
    bool operands_available(int32_t L) const {
        const Node& nd = rt.nodes[info[L].node];
        for (int k = 0; k < 3; k++) {
            if (nd.in[k] < 0) continue;
            if (!available(rt.nodes[nd.in[k]].local)) return false;
        }
        return true;
    }

    // end the current kernel early: everything live in acc / registers that is still needed goes to HBM
    void cut() {
        if (acc_owner >= 0 && rem(acc_owner) > 0 && !has_buf(acc_owner)) store_value(acc_owner);
        for (int j = 0; j < TAPE_REGS; j++) {
            const int32_t L = reg_owner[j];
            if (L >= 0 && rem(L) > 0 && !has_buf(L)) store_value(L);
        }
        launch(RM_NONE, 0.0, -1);
        begin_kernel();
    }

    // ---- peephole fusion on the abstract code: every dispatch costs ~18 issue slots, so fewer, fatter instructions ----
    //   MUL_I a ; ADD_I b                      -> MULADD_II a, b          (two roundings, like the pair)
    //   ADD_I a | SUB_I a ; MUL_I b             -> ADDMUL_II +-a, b
    //   ADD_S r ; STR r                         -> ACCUM_S r               (running sum kept in the register file)
    //   STR t ; MOV x ; MUL_I a ; ADD_S t       -> ADDPROD x, a            (t dead afterwards; float add commutes exactly)
    //   (acc -> t) ; MOV x ; MUL_I a ; ADD_I 1 ; VID_S t  -> DISCOUNT x, a     acc / (1 + x * a), the same three roundings
    //   (acc -> t) ; MOV x ; MUL_I a ; ADD_I 1 ; MUL_S t  -> ACCRUE x, a       acc * (1 + x * a)
    bool reg_dead_after(size_t from, int reg) const {
        for (size_t i = from; i < A.size(); i++) {
            const AIns& a = A[i];
            if (a.op == T_STR && a.kind == K_REG && a.arg == reg) return true;        // overwritten before any read
            if (a.kind == K_REG && a.arg == reg) return false;                        // read (binary operand, STGS, ACCUM, END)
        }
        return true;
    }
    void peephole() {
        std::vector<AIns> out;
        out.reserve(A.size());
        int leaf_refs = 0;
        for (size_t i = 0; i < A.size(); i++) {
            const AIns& a = A[i];
            // STR t ; MOV x ; ADD_I|SUB_I a ; MUL_I b ; ADD_S t   -> ADDAFF x, +-a, b     (t dead afterwards)
            if (i + 4 < A.size() && a.op == T_STR && a.kind == K_REG
                && A[i + 1].op == (A_BIN | B_MOV) && (A[i + 1].kind == K_REG || A[i + 1].kind == K_LEAF)
                && !(A[i + 1].kind == K_REG && A[i + 1].arg == a.arg)
                && (A[i + 2].op == (A_BIN | B_ADD) || A[i + 2].op == (A_BIN | B_SUB)) && A[i + 2].kind == K_IMM
                && A[i + 3].op == (A_BIN | B_MUL) && A[i + 3].kind == K_IMM
                && A[i + 4].op == (A_BIN | B_ADD) && A[i + 4].kind == K_REG && A[i + 4].arg == a.arg
                && reg_dead_after(i + 5, a.arg)) {
                const bool sub = A[i + 2].op == (A_BIN | B_SUB);
                const uint32_t first = sub ? (A[i + 2].y ^ 0x80000000u) : A[i + 2].y;
                // y = first immediate, y of the following pseudo-entry = second immediate (see schedule)
                out.push_back(mk((uint16_t)T_ADDAFF_S, A[i + 1].kind, A[i + 1].arg, first, A[i + 2].src, sub ? 1 : 0));
                out.push_back(mk(A_EXT, K_NONE, 0, A[i + 3].y, A[i + 3].src));
                if (A[i + 1].kind == K_LEAF) leaf_refs++;
                i += 4;
                continue;
            }
            if (i + 3 < A.size() && a.op == T_STR && a.kind == K_REG
                && A[i + 1].op == (A_BIN | B_MOV) && (A[i + 1].kind == K_REG || A[i + 1].kind == K_LEAF)
                && !(A[i + 1].kind == K_REG && A[i + 1].arg == a.arg)
                && A[i + 2].op == (A_BIN | B_MUL) && A[i + 2].kind == K_IMM
                && A[i + 3].op == (A_BIN | B_ADD) && A[i + 3].kind == K_REG && A[i + 3].arg == a.arg
                && reg_dead_after(i + 4, a.arg)) {
                out.push_back(mk((uint16_t)(A_BIN | B_ADDPROD), A[i + 1].kind, A[i + 1].arg, A[i + 2].y, A[i + 2].src));
                if (A[i + 1].kind == K_LEAF) leaf_refs++;
                i += 3;
                continue;
            }
            // (acc -> t) ; MOV x ; MUL_I a ; ADD_I 1 ; VID_S t | MUL_S t     -> DISCOUNT x, a | ACCRUE x, a   on acc   (t dead afterwards)
            if (i + 3 < A.size() && !out.empty() && a.op == (A_BIN | B_MOV) && (a.kind == K_REG || a.kind == K_LEAF)
                && A[i + 1].op == (A_BIN | B_MUL) && A[i + 1].kind == K_IMM
                && A[i + 2].op == (A_BIN | B_ADD) && A[i + 2].kind == K_IMM && A[i + 2].y == f2u(1.0f)
                && (A[i + 3].op == (A_BIN | B_VID) || A[i + 3].op == (A_BIN | B_MUL)) && A[i + 3].kind == K_REG
                && (out.back().op == T_STR || out.back().op == T_ACCUM_S) && out.back().kind == K_REG && out.back().arg == A[i + 3].arg
                && !(a.kind == K_REG && a.arg == A[i + 3].arg) && reg_dead_after(i + 4, A[i + 3].arg)) {
                const int t = A[i + 3].arg;
                if (out.back().op == T_STR) out.pop_back();                                   // acc still holds the value
                else out.back() = mk((uint16_t)(A_BIN | B_ADD), K_REG, t, 0u);              // ACCUM without the write-back
                out.push_back(mk((uint16_t)(A_BIN | (A[i + 3].op == (A_BIN | B_VID) ? B_DISCOUNT : B_ACCRUE)), a.kind, a.arg, A[i + 1].y, A[i + 1].src));
                if (a.kind == K_LEAF) leaf_refs++;
                i += 3;
                continue;
            }
            // x * 1.0f and x / 1.0f are x for every float (NaN stays NaN, -0 stays -0): no instruction (a swaption value is multiplied
            // by the numeraire at time 0, the scalar 1). The tape-cache key records which scalar operands equal 1.0f.
            if ((a.op == (A_BIN | B_MUL) || a.op == (A_BIN | B_DIV)) && a.kind == K_IMM && a.y == f2u(1.0f)) continue;
            if (i + 1 < A.size() && a.op == (A_BIN | B_MUL) && a.kind == K_IMM && A[i + 1].op == (A_BIN | B_ADD) && A[i + 1].kind == K_IMM) {
                out.push_back(mk((uint16_t)T_MULADD_II, K_IMM, (int32_t)A[i + 1].y, a.y, a.src, 0, A[i + 1].src));   // arg carries the second immediate's bits
                i += 1;
                continue;
            }
            if (i + 1 < A.size() && (a.op == (A_BIN | B_ADD) || a.op == (A_BIN | B_SUB)) && a.kind == K_IMM && A[i + 1].op == (A_BIN | B_MUL) && A[i + 1].kind == K_IMM) {
                // x - a == x + (-a) bit for bit
                const bool sub = a.op == (A_BIN | B_SUB);
                const uint32_t first = sub ? (a.y ^ 0x80000000u) : a.y;
                out.push_back(mk((uint16_t)T_ADDMUL_II, K_IMM, (int32_t)A[i + 1].y, first, a.src, sub ? 1 : 0, A[i + 1].src));
                i += 1;
                continue;
            }
            // (not when the STR opens an ADDPROD: "ADD_S r ; STR r ; MOV x ; MUL_I a ; ADD_S r" is ADD_S r ; ADDPROD x, a)
            if (i + 1 < A.size() && a.op == (A_BIN | B_ADD) && a.kind == K_REG && A[i + 1].op == T_STR && A[i + 1].kind == K_REG && A[i + 1].arg == a.arg
                && !(i + 4 < A.size() && A[i + 2].op == (A_BIN | B_MOV) && (A[i + 2].kind == K_REG || A[i + 2].kind == K_LEAF)
                     && !(A[i + 2].kind == K_REG && A[i + 2].arg == a.arg)
                     && A[i + 3].op == (A_BIN | B_MUL) && A[i + 3].kind == K_IMM
                     && A[i + 4].op == (A_BIN | B_ADD) && A[i + 4].kind == K_REG && A[i + 4].arg == a.arg && reg_dead_after(i + 5, a.arg))) {
                out.push_back(mk((uint16_t)T_ACCUM_S, K_REG, a.arg, 0u));
                i += 1;
                continue;
            }
            if (a.kind == K_LEAF) leaf_refs++;
            out.push_back(a);
        }
        // second pass, over the fused forms: whole LMM drift terms and swaption periods in one dispatch
        //   MULADD_II a, b ; VID_I c ; MUL_I d          -> RATIO a, b, c, d        (c / (acc * a + b)) * d
        //   MULADD_II a, b ; MUL_I c                    -> MULADDMUL a, b, c
        //   ADDAFF x, a, b ; DISCOUNT x, p              -> ADDAFFDISC x, a, b, p   (same operand x)
        A.clear();
        for (size_t i = 0; i < out.size(); i++) {
            const AIns& a = out[i];
            if (a.op == T_MULADD_II && i + 2 < out.size() && out[i + 1].op == (A_BIN | B_VID) && out[i + 1].kind == K_IMM
                && out[i + 2].op == (A_BIN | B_MUL) && out[i + 2].kind == K_IMM) {
                A.push_back(mk((uint16_t)T_RATIO, K_IMM, a.arg, a.y, a.src, 0, a.src2));
                A.push_back(mk(A_EXT, K_NONE, 0, out[i + 1].y, out[i + 1].src));
                A.push_back(mk(A_EXT, K_NONE, 0, out[i + 2].y, out[i + 2].src));
                i += 2;
                continue;
            }
            if (a.op == T_MULADD_II && i + 1 < out.size() && out[i + 1].op == (A_BIN | B_MUL) && out[i + 1].kind == K_IMM) {
                A.push_back(mk((uint16_t)T_MULADDMUL, K_IMM, a.arg, a.y, a.src, 0, a.src2));
                A.push_back(mk(A_EXT, K_NONE, 0, out[i + 1].y, out[i + 1].src));
                i += 1;
                continue;
            }
            if (a.op == T_ADDAFF_S && i + 2 < out.size() && out[i + 1].op == A_EXT && out[i + 2].op == (A_BIN | B_DISCOUNT)
                && out[i + 2].kind == a.kind && out[i + 2].arg == a.arg && (a.kind == K_REG || a.kind == K_LEAF)) {
                A.push_back(mk((uint16_t)T_ADDAFFDISC_S, a.kind, a.arg, a.y, a.src, a.neg));
                A.push_back(out[i + 1]);
                A.push_back(mk(A_EXT, K_NONE, 0, out[i + 2].y, out[i + 2].src));
                if (a.kind == K_LEAF) leaf_refs--;
                i += 2;
                continue;
            }
            A.push_back(a);
        }
        // third pass: one dispatch per LMM drift term / state update (at ~1 chunk per warp the interpreter runs at the latency of
        // a lone warp: ~130 cycles per dispatch whatever the handler does, so the number of dispatches is what counts)
        //   MOV x ; RATIO a, b, c, d ; ACCUM_S r                        -> RATIOACC x, a, b, c, d | r
        //   MULADDMUL a, b, c ; ADD x ; ADDPROD w, d ; STG p            -> AXPYST x, a, b, c, d | w, p, (reload)
        if (rt.opt.fuse_ops2) {
            out.swap(A);
            A.clear();
            auto is_src = [](const AIns& t) { return t.kind == K_REG || t.kind == K_LEAF; };
            for (size_t i = 0; i < out.size(); i++) {
                const AIns& a = out[i];
                if (a.op == (A_BIN | B_MOV) && is_src(a) && i + 4 < out.size() && out[i + 1].op == T_RATIO && out[i + 2].op == A_EXT && out[i + 3].op == A_EXT
                    && out[i + 4].op == T_ACCUM_S && out[i + 4].kind == K_REG && !(a.kind == K_REG && a.arg == out[i + 4].arg)) {
                    const AIns& r = out[i + 1];
                    A.push_back(mk((uint16_t)T_RATIOACC_S, a.kind, a.arg, r.y, r.src, 0));
                    A.push_back(mk(A_EXT, K_NONE, 0, (uint32_t)r.arg, r.src2));
                    A.push_back(out[i + 2]);
                    A.push_back(mk(A_EXT, K_REG, out[i + 4].arg, out[i + 3].y, out[i + 3].src));
                    i += 4;
                    continue;
                }
                //   STR t ; RATIO a, b, c, d ; ACCUM_S r                        -> RATIOACC_A t, a, b, c, d | r   (the state is in acc)
                if (a.op == T_STR && a.kind == K_REG && i + 4 < out.size() && out[i + 1].op == T_RATIO && out[i + 2].op == A_EXT && out[i + 3].op == A_EXT
                    && out[i + 4].op == T_ACCUM_S && out[i + 4].kind == K_REG && a.arg != out[i + 4].arg) {
                    const AIns& r = out[i + 1];
                    A.push_back(mk((uint16_t)T_RATIOACC_A, K_REG, a.arg, r.y, r.src, 0));
                    A.push_back(mk(A_EXT, K_NONE, 0, (uint32_t)r.arg, r.src2));
                    A.push_back(out[i + 2]);
                    A.push_back(mk(A_EXT, K_REG, out[i + 4].arg, out[i + 3].y, out[i + 3].src));
                    i += 4;
                    continue;
                }
                if (a.op == T_MULADDMUL && i + 4 < out.size() && out[i + 1].op == A_EXT && out[i + 2].op == (A_BIN | B_ADD) && is_src(out[i + 2])
                    && out[i + 3].op == (A_BIN | B_ADDPROD) && is_src(out[i + 3]) && out[i + 4].op == T_STG
                    && !(out[i + 2].kind == out[i + 3].kind && out[i + 2].arg == out[i + 3].arg)) {
                    const AIns& x = out[i + 2];
                    const AIns& w = out[i + 3];
                    A.push_back(mk((uint16_t)T_AXPYST_S, x.kind, x.arg, a.y, a.src, 0));
                    A.push_back(mk(A_EXT, K_NONE, 0, (uint32_t)a.arg, a.src2));
                    A.push_back(out[i + 1]);
                    A.push_back(mk(A_EXT, w.kind, w.arg, w.y, w.src));
                    A.push_back(mk(A_EXT, K_NONE, 0, out[i + 4].y, -1));
                    A.push_back(mk(A_EXT, K_RELOAD, 0, 0xffffffffu, -1));
                    i += 4;
                    continue;
                }
                A.push_back(a);
            }
        }
        n_leaf_refs = leaf_refs;
    }

    // ---- ring scheduling: abstract code -> tape (see tape_isa.h) ----
    struct Event { int32_t leaf; std::vector<int32_t> use; size_t k = 0; int slot = -1; bool waited = false; };

    // patches: where the immediates of the abstract code ended up (indices into `body`)
    void schedule(int ring_max, bool pipeline, int horizon, int shift, std::vector<TapeInstr>& prologue, std::vector<TapeInstr>& body, int& n_ring,
                  std::vector<ImmPatch>& patches) {
        auto note = [&](const AIns& a) { if (a.src >= 0) patches.push_back(ImmPatch{(int32_t)body.size() - 1, a.src, a.neg}); };
        // 1. residency intervals ("events") of every leaf: consecutive uses closer than `horizon` share one TMA copy
        std::vector<Event> ev;
        {
            std::map<int32_t, int> open;    // leaf -> index of its latest event
            for (int32_t i = 0; i < (int32_t)A.size(); i++) {
                if (A[i].kind != K_LEAF) continue;
                const int32_t L = A[i].arg;
                auto it = open.find(L);
                if (it != open.end() && i - ev[it->second].use.back() <= horizon) ev[it->second].use.push_back(i);
                else { Event e; e.leaf = L; e.use.push_back(i); open[L] = (int)ev.size(); ev.push_back(std::move(e)); }
            }
        }
        // events are created in order of their first use; `order` is the queue of events still to be issued
        std::vector<int> order(ev.size());
        for (size_t k = 0; k < ev.size(); k++) order[k] = (int)k;
        size_t q = 0;
        n_ring = (int)std::min<size_t>((size_t)ring_max, ev.size());
        const uint32_t R = (uint32_t)n_ring;
        std::vector<int> slot_ev(R, -1), pro_leaf(R, -1);
        std::vector<char> loadn_done(R, 0);
        std::vector<int> ev_at(A.size(), -1);   // instruction -> event it reads
        auto map_uses = [&](int e) { for (size_t k = ev[e].k; k < ev[e].use.size(); k++) ev_at[ev[e].use[k]] = e; };
        for (size_t k = 0; k < ev.size(); k++) map_uses((int)k);

        auto issue = [&](std::vector<TapeInstr>& out, int e, int s) {
            ev[e].slot = s; ev[e].waited = false; slot_ev[s] = e;
            out.push_back(enc_idx(T_LOAD, (uint32_t)s, (uint32_t)info[ev[e].leaf].slot, shift));
        };
        auto refill = [&](int s) {
            slot_ev[s] = -1;
            if (q < order.size()) issue(body, order[q++], s);
            else if (pipeline && pro_leaf[s] >= 0 && !loadn_done[s]) {
                body.push_back(enc_idx(T_LOADN, (uint32_t)s, (uint32_t)info[pro_leaf[s]].slot, shift));
                loadn_done[s] = 1;
            }
        };
        // 2. fill the ring: before the first chunk (prologue) when pipelining, else at the top of the body
        for (uint32_t s = 0; s < R && q < order.size(); s++) {
            const int e = order[q++];
            issue(pipeline ? prologue : body, e, (int)s);
            if (pipeline) pro_leaf[s] = ev[e].leaf;
        }
        // 3. walk the code
        // make sure event e sits in a ring slot (its T_LOAD has been emitted)
        auto ensure_issued = [&](int e) {
            if (ev[e].slot >= 0) return;
            // not issued yet and no slot was free in time: take the slot whose occupant is needed last
            // (while an event is waiting in the queue every slot is occupied: freed slots are refilled at once)
            int vs = -1; int32_t far = -1;
            for (uint32_t s = 0; s < R; s++) {
                const int o = slot_ev[s];
                if (o < 0) continue;
                const int32_t nu = ev[o].use[ev[o].k];
                if (nu > far) { far = nu; vs = (int)s; }
            }
            if (vs < 0) fail(FMC_ERR_UNSUPPORTED, "internal: no ring slot to evict");
            const int o = slot_ev[vs];
            {
                if (!ev[o].waited) body.push_back(enc_idx(T_WAIT, (uint32_t)vs, 0, shift));
                // the evicted occupant's remaining uses become a new event, queued by its next use
                Event rest; rest.leaf = ev[o].leaf;
                rest.use.assign(ev[o].use.begin() + (long)ev[o].k, ev[o].use.end());
                ev[o].use.resize(ev[o].k);
                const int ne = (int)ev.size();
                ev.push_back(std::move(rest));
                map_uses(ne);
                size_t at = q;
                while (at < order.size() && ev[order[at]].use[0] < ev[ne].use[0]) at++;
                order.insert(order.begin() + (long)at, ne);
            }
            // e is somewhere in the queue (normally its head): take it out
            for (size_t k = q; k < order.size(); k++) if (order[k] == e) { order.erase(order.begin() + (long)k); break; }
            issue(body, e, vs);
        };
        // what has to happen once the extension words of the instruction being emitted are out: ring slots whose last use it
        // was are refilled (their T_LOAD must not separate an instruction from its extension words)
        struct Pending { int ext_left; int slot; bool append; uint32_t value; };
        std::vector<Pending> pending;
        uint32_t fused_reload = 0xffffffffu;     // pointer-table slot for the K_RELOAD word of the instruction being emitted
        auto after_word = [&]() {
            for (size_t k = 0; k < pending.size();) {
                if (--pending[k].ext_left > 0) { k++; continue; }
                const Pending pd = pending[k];
                pending.erase(pending.begin() + (long)k);
                if (pd.append) body.push_back(TapeInstr{ T_END, pd.value });
                else refill(pd.slot);
            }
        };
        // the slot's last use is a form that can re-arm it itself: take the next event off the queue without emitting a T_LOAD
        auto take_reload = [&](int s, uint32_t& ptr) -> bool {
            if (q >= order.size()) return false;
            const int e = order[q++];
            ev[e].slot = s; ev[e].waited = false; slot_ev[s] = e;
            ptr = (uint32_t)info[ev[e].leaf].slot;
            return true;
        };
        for (int32_t i = 0; i < (int32_t)A.size(); i++) {
            const AIns& a = A[i];
            int n_ext = 0;
            if (a.op != A_EXT) while (i + 1 + n_ext < (int32_t)A.size() && A[i + 1 + n_ext].op == A_EXT) n_ext++;
            if (a.op == A_EXT) {
                uint32_t x = T_END, y = a.y;
                if (a.kind == K_REG) x |= (R + (uint32_t)a.arg) << shift;
                else if (a.kind == K_RELOAD) y = fused_reload;
                else if (a.kind == K_LEAF) {
                    Event& E2 = ev[ev_at[i]];
                    x |= (uint32_t)E2.slot << shift;
                    E2.k++;
                    if (E2.k >= E2.use.size()) {
                        int left = 1;                                      // this word and the extension words still to come
                        for (int32_t j = i + 1; j < (int32_t)A.size() && A[j].op == A_EXT; j++) left++;
                        pending.push_back(Pending{left, E2.slot, false, 0u});
                    }
                }
                body.push_back(TapeInstr{ x, y });
                note(a);
                after_word();
            } else if (a.kind == K_LEAF) {
                const int e = ev_at[i];
                ensure_issued(e);
                // second operands of the fused forms travel in extension words: in their slot and waited for before the instruction
                for (int32_t j = i + 1; j <= i + n_ext; j++) {
                    if (A[j].kind != K_LEAF) continue;
                    const int e2 = ev_at[j];
                    ensure_issued(e2);
                    if (!ev[e2].waited) { body.push_back(enc_idx(T_WAIT, (uint32_t)ev[e2].slot, 0, shift)); ev[e2].waited = true; }
                }
                if (ev_at[i] != e || ev[e].slot < 0) fail(FMC_ERR_UNSUPPORTED, "internal: operand evicted by the second operand of its own instruction");
                Event& E = ev[e];
                if (a.op == T_AXPYST_S && !E.waited) { body.push_back(enc_idx(T_WAIT, (uint32_t)E.slot, 0, shift)); E.waited = true; }
                const uint32_t fl = E.waited ? 1u : 2u;        // _S / _W
                E.waited = true;
                const bool last = E.k + 1 >= E.use.size();
                uint32_t opc, reload_ptr = 0xffffffffu;
                bool append_reload = false;
                fused_reload = 0xffffffffu;
                if (a.op == T_ADDAFF_S) opc = fl == 1u ? (uint32_t)T_ADDAFF_S : (uint32_t)T_ADDAFF_W;
                else if (a.op == T_ADDAFFDISC_S) {
                    if (last && rt.opt.fuse_ops2 && take_reload(E.slot, reload_ptr)) { opc = fl == 1u ? (uint32_t)T_ADDAFFDISC_SL : (uint32_t)T_ADDAFFDISC_WL; append_reload = true; }
                    else opc = fl == 1u ? (uint32_t)T_ADDAFFDISC_S : (uint32_t)T_ADDAFFDISC_W;
                }
                else if (a.op == T_RATIOACC_S) opc = fl == 1u ? (uint32_t)T_RATIOACC_S : (uint32_t)T_RATIOACC_W;
                else if (a.op == T_AXPYST_S) { opc = (uint32_t)T_AXPYST_S; if (last && take_reload(E.slot, reload_ptr)) fused_reload = reload_ptr; }
                else opc = T_BIN0 + 3u * (uint32_t)(a.op & 0xff) + fl;
                const int slot_used = E.slot;
                body.push_back(TapeInstr{ opc | ((uint32_t)slot_used << shift), a.y });
                note(a);
                ev[e].k++;
                // a multi-word instruction keeps its extension words right behind it: the slot is refilled after them
                if (last) {
                    if (append_reload) pending.push_back(Pending{n_ext, slot_used, true, reload_ptr});
                    else if (reload_ptr != 0xffffffffu) { /* re-armed by the instruction itself */ }
                    else if (n_ext > 0) pending.push_back(Pending{n_ext, slot_used, false, 0u});
                    else refill(slot_used);
                }
            } else if (a.op & A_BIN) {
                const uint32_t bop = T_BIN0 + 3u * (uint32_t)(a.op & 0xff);
                if (a.kind == K_IMM) body.push_back(TapeInstr{ bop, a.y });
                else body.push_back(TapeInstr{ (bop + 1u) | ((R + (uint32_t)a.arg) << shift), a.y });
                note(a);
            } else if (a.op == T_MULADD_II || a.op == T_ADDMUL_II || a.op == T_MULADDMUL || a.op == T_RATIO) {
                body.push_back(TapeInstr{ (uint32_t)a.op, a.y });
                note(a);
                body.push_back(TapeInstr{ T_END, (uint32_t)a.arg });      // extension word: only its y is read
                if (a.src2 >= 0) patches.push_back(ImmPatch{(int32_t)body.size() - 1, a.src2, 0});
            } else if (a.op == T_END) {
                // re-arm whatever prologue slot has not been re-armed yet (only slots that were never freed: none in practice)
                if (a.kind == K_REG) body.push_back(TapeInstr{ T_END | ((R + (uint32_t)a.arg) << shift), 1u });
                else body.push_back(TapeInstr{ T_END, 0u });
            } else {
                // (the fused forms with their first operand in the register file land here too: _S with a register-file slot)
                for (int32_t j = i + 1; j <= i + n_ext; j++) {
                    if (A[j].kind != K_LEAF) continue;
                    const int e2 = ev_at[j];
                    ensure_issued(e2);
                    if (!ev[e2].waited) { body.push_back(enc_idx(T_WAIT, (uint32_t)ev[e2].slot, 0, shift)); ev[e2].waited = true; }
                }
                fused_reload = 0xffffffffu;
                const uint32_t slot = (a.kind == K_REG) ? R + (uint32_t)a.arg : 0u;
                body.push_back(TapeInstr{ (uint32_t)a.op | (slot << shift), a.y });
                if (a.op != T_STG && a.op != T_STGS) note(a);
            }
        }
        if (!pending.empty()) fail(FMC_ERR_UNSUPPORTED, "internal: extension words missing behind a multi-word instruction");
        if (pipeline) for (uint32_t s = 0; s < R; s++)
            if (pro_leaf[s] >= 0 && !loadn_done[s]) fail(FMC_ERR_UNSUPPORTED, "internal: ring slot %u not re-armed", s);
    }

    // FMC_DUMP_TAPES=1: the abstract code of every kernel after fusion, one instruction per line (development aid)
    void dump_abstract() const {
        static const char* const bin[] = {"MOV", "ADD", "SUB", "BUS", "MUL", "DIV", "VID", "MIN", "MAX", "SEL", "ADDPROD", "ACCRUE", "DISCOUNT"};
        auto name = [&](uint16_t op) -> std::string {
            if (op == A_EXT) return "  ext";
            if (op & A_BIN) return bin[op & 0xff];
            switch (op) {
            case T_END: return "END"; case T_STG: return "STG"; case T_STGS: return "STGS"; case T_STR: return "STR"; case T_SETP: return "SETP";
            case T_SQR: return "SQR"; case T_MULADD_II: return "MULADD_II"; case T_ACCUM_S: return "ACCUM_S"; case T_ADDMUL_II: return "ADDMUL_II";
            case T_ADDAFF_S: return "ADDAFF"; case T_MULADDMUL: return "MULADDMUL"; case T_RATIO: return "RATIO"; case T_ADDAFFDISC_S: return "ADDAFFDISC";
            case T_RATIOACC_S: return "RATIOACC"; case T_AXPYST_S: return "AXPYST"; case T_RATIOACC_A: return "RATIOACC_A";
            default: return "op" + std::to_string(op);
            }
        };
        std::fprintf(stderr, "[fmc dump] kernel: %zu abstract instructions\n", A.size());
        for (const AIns& a : A) {
            std::string arg;
            if (a.kind == K_REG) arg = "r" + std::to_string(a.arg);
            else if (a.kind == K_LEAF) arg = "leaf" + std::to_string(a.arg) + (info[a.arg].lazy ? "*" : "");
            else if (a.kind == K_IMM) arg = "imm";
            else if (a.kind == K_RELOAD) arg = "reload";
            std::fprintf(stderr, "[fmc dump]   %-12s %s%s\n", name(a.op).c_str(), arg.c_str(), (a.op == T_STG || a.op == T_STGS) ? (" -> p" + std::to_string(a.y)).c_str() : "");
        }
    }

    std::vector<KernelPlan> plans;      // the launches of this cone, in order (kept for the tape cache)
    bool keep_plans = false;

    void launch(int reduce_mode, double reduce_param, int32_t weight_local) {
        if (reduce_mode == RM_DOT || reduce_mode == RM_WSQ) {
            // epilogue convention: the VALUE waits in a register-file slot, the WEIGHT sits in acc (so that the ring
            // slot of the weight leaf is free to be re-armed before T_END)
            const int32_t V = acc_owner;
            set_pins(V, weight_local);
            if (info[V].reg < 0) put_acc_in_reg();
            const int vreg = info[V].reg;
            emit_src(B_MOV, weight_local);
            emit(T_END, K_REG, vreg);
        } else emit(T_END);
        if (A.size() <= 1 && reduce_mode == RM_NONE) return;   // nothing to do
        { PhaseTimer pt(3); if (rt.opt.fuse_ops) peephole(); }

        static const bool dump_tapes = std::getenv("FMC_DUMP_TAPES") != nullptr;
        if (dump_tapes) dump_abstract();
        std::vector<TapeInstr> prologue, body;
        KernelPlan kp;
        int n_ring = 0;
        int n_warps = std::max(1, std::min(rt.windowing && rt.opt.window_cta_warps > 0 ? rt.opt.window_cta_warps : rt.opt.cta_warps, TAPE_MAX_WARPS));
        // Chunk geometry (elements per lane, tape_interp.cuh). A warp interprets one chunk of 32 E paths at a time, so a vector
        // of n paths is n / (32 E) warps' worth of work per pass: the largest E that still gives every SM `min_warps` warps
        // (16-element chunks amortise the dispatch best; below that the GPU's warp slots stay empty and the interpreter runs
        // at the latency of a lone warp, so more, shorter warps win).
        int elems = rt.opt.tape_elems;
        // a window kernel carries a register-file slot per level and chain state: halving the slot size doubles the warps that fit
        if (!tape_valid_elems(elems) && rt.windowing && tape_valid_elems(rt.opt.window_elems) && n >= (int64_t)rt.opt.min_warps * rt.sm_count * tape_chunk(rt.opt.window_elems))
            elems = rt.opt.window_elems;
        if (!tape_valid_elems(elems)) {
            elems = 4;
            for (int e = TAPE_E_MAX; e > 4; e >>= 1)
                if (n >= (int64_t)rt.opt.min_warps * rt.sm_count * tape_chunk(e)) { elems = e; break; }
        }
        n_warps = std::min(n_warps, elems == 16 ? 16 : elems == 8 ? 28 : 32);       // 64 K registers per CTA at 128 / 72 / 40 per thread
        const int slot_bytes = tape_slot_bytes(elems);
        const int64_t chunks = (n + tape_chunk(elems) - 1) / tape_chunk(elems);
        // Shared-memory budget of one warp, in slots, if `target` CTAs are to be resident per SM: as many as the vector needs
        // to be resident at once (at ~1 chunk per warp a grid that needs a second round of CTAs runs it at a fraction of the
        // occupancy), at most what the registers of this geometry allow; short tapes keep the occupancy high, long ones
        // trade it for ring depth (never below ring_min slots).
        const size_t est_tables = 8 * (A.size() + 2 * (size_t)n_leaf_refs + 2 * TAPE_MAX_RING + 6) + 8 * ptrs.size() + 256;
        static OccCache occ;
        auto blocks_per_sm = [&](size_t smem_bytes) {
            const auto key = std::make_pair((smem_bytes + 1023) / 1024, (reduce_mode * 16 + n_warps) * 32 + elems);
            auto it = occ.blocks.find(key);
            if (it == occ.blocks.end()) it = occ.blocks.emplace(key, tape_max_blocks_per_sm(key.first * 1024, reduce_mode, n_warps, elems)).first;
            return it->second;
        };
        const int hw_max = blocks_per_sm(1024);
        const int64_t need = (chunks + (int64_t)n_warps * rt.sm_count - 1) / ((int64_t)n_warps * rt.sm_count);
        int target = (int)std::max<int64_t>(1, std::min<int64_t>(need, hw_max));
        if (rt.opt.target_ctas > 0) target = std::min(rt.opt.target_ctas, hw_max);          // tests / tuning: forced
        int ring_want = TAPE_MAX_RING;
        auto slots_for = [&](int ctas) {
            const size_t cta_share = rt.smem_per_sm / (size_t)ctas;
            const long budget_bytes = (long)std::min(cta_share, rt.smem_per_cta_max) - 1024 - (ctas > 1 ? 1024 : 0) - (long)est_tables;
            return (int)std::max<long>(1, budget_bytes / (n_warps * slot_bytes));
        };
        if (rt.windowing && rt.opt.target_ctas <= 0) {
            // a window keeps some leaves in the ring for the whole kernel (the Brownian increment of each of its time steps is read
            // once per component): the ring needs a slot for each of them plus a few for the leaves that stream through, or every
            // access becomes a fresh copy. Occupancy gives way until that fits.
            static thread_local std::vector<int32_t> cnt_leaf;
            cnt_leaf.clear();
            int n_hot = 0;
            for (const AIns& a : A) if (a.kind == K_LEAF) {
                if ((size_t)a.arg >= cnt_leaf.size()) cnt_leaf.resize((size_t)a.arg + 1, 0);
                if (++cnt_leaf[(size_t)a.arg] == 4) n_hot++;
            }
            ring_want = std::min(n_hot + rt.opt.window_ring_extra, TAPE_MAX_RING);
            while (target > 1 && slots_for(target) - regs_used < ring_want) target--;
        }
        const int slot_budget = slots_for(target);
        int ring_max = std::max(1, std::min<int>(rt.opt.ring_max, TAPE_MAX_RING));
        ring_max = std::min(ring_max, std::max(rt.opt.ring_min, std::min(slot_budget - regs_used, ring_want)));
        { PhaseTimer pt(4); schedule(ring_max, rt.opt.pipeline, rt.opt.horizon, tape_slot_shift(elems), prologue, body, n_ring, kp.patches); }
        PhaseTimer pt_prep(5);
        const size_t total = prologue.size() + 1 + body.size();
        if (total > (size_t)TAPE_MAX_INSTR + 1 || ptrs.size() > (size_t)TAPE_MAX_PTRS)
            fail(FMC_ERR_UNSUPPORTED, "internal: tape overflow (%zu instr, %zu ptrs)", total, ptrs.size());
        kp.n_instr = (int)total;
        kp.n_prologue = (int)prologue.size();
        kp.n_ring = n_ring;
        kp.n_slots = n_ring + regs_used;
        kp.reduce_mode = reduce_mode;
        kp.n_warps = n_warps;
        kp.elems = elems;
        kp.n_leaf_slots = n_leaf_slots; kp.n_result_stores = n_result_stores;
        kp.words.reserve(total + 2);
        kp.words.insert(kp.words.end(), prologue.begin(), prologue.end());
        kp.words.push_back(TapeInstr{ T_END, 0u });             // closes the prologue
        kp.words.insert(kp.words.end(), body.begin(), body.end());
        kp.words.push_back(TapeInstr{ T_END, 0u });             // the interpreter prefetches two words ahead
        kp.words.push_back(TapeInstr{ T_END, 0u });
        for (ImmPatch& ip : kp.patches) ip.word += (int32_t)prologue.size() + 1;
        kp.ptr_local = slotted;
        // slot sets: a tape that leaves most of the budget unused keeps several chunks per warp in flight
        int n_sets = 1, per_sm = 1, grid = 1;
        size_t smem = 0;
        if (rt.opt.pipeline && n_ring > 0) n_sets = std::max(1, std::min(rt.opt.max_sets, slot_budget / std::max(1, kp.n_slots)));
        for (;;) {
            smem = tape_smem_bytes((int)ptrs.size(), kp.n_instr, kp.n_slots, n_sets, n_warps, elems);
            if (smem > rt.smem_per_cta_max && n_sets > 1) { n_sets--; continue; }
            if (smem > rt.smem_per_cta_max && n_warps > 1) { n_warps = std::max(1, n_warps / 2); kp.n_warps = n_warps; continue; }   // a wide CTA of a forced geometry
            if (smem > rt.smem_per_cta_max) fail(FMC_ERR_UNSUPPORTED, "internal: tape needs %zu bytes of shared memory per CTA", smem);
            per_sm = blocks_per_sm(smem);
            grid = (int)std::min<int64_t>((chunks + n_warps - 1) / n_warps, (int64_t)per_sm * rt.sm_count);
            grid = std::max(1, std::min(grid, rt.max_grid));
            if (rt.opt.grid_limit > 0) grid = std::min(grid, rt.opt.grid_limit);
            const int64_t chunks_per_warp = (chunks + (int64_t)grid * n_warps - 1) / ((int64_t)grid * n_warps);
            if (n_sets > 1 && n_sets > chunks_per_warp) { n_sets = (int)std::max<int64_t>(1, chunks_per_warp); continue; }
            break;
        }
        kp.n_sets = n_sets;
        kp.grid = grid;
        static const bool log_tapes = std::getenv("FMC_LOG_TAPES") != nullptr;
        if (log_tapes)
            std::fprintf(stderr, "[fmc tape] n=%lld elems=%d instr=%zu (abstract %zu, prologue %zu) ptrs=%zu leaves=%d stores=%d ring=%d regs=%d sets=%d warps=%d smem=%zu ctas/sm=%d grid=%d reduce=%d\n",
                         (long long)n, elems, total, A.size(), prologue.size(), ptrs.size(), n_leaf_slots, n_result_stores, n_ring, regs_used, n_sets, n_warps, smem, per_sm, grid, reduce_mode);
        submit(kp, reduce_param, false);
        if (keep_plans) plans.push_back(std::move(kp));
    }

    // fill the launch parameters from a plan and the buffers / immediates of THIS cone, launch
    void submit(const KernelPlan& kp, double reduce_param, bool patch) {
        TapeParams& P = *params;
        P.n = n;
        P.elems = kp.elems;
        P.n_instr = kp.n_instr;
        P.n_prologue = kp.n_prologue;
        P.n_ptrs = (int)kp.ptr_local.size();
        P.n_ring = kp.n_ring;
        P.n_slots = kp.n_slots;
        P.n_sets = kp.n_sets;
        P.reduce_mode = kp.reduce_mode;
        P.reduce_param = reduce_param;
        P.partials = rt.d_partials;
        P.counter = rt.d_counter;
        P.result = rt.d_result;
        P.host_result = nullptr; P.ticket = 0.0;
        for (int r = 0; r < XMAX_RANKS; r++) P.xchg.tables[r] = nullptr;
        P.xchg.host_table = nullptr;
        P.xchg.rank = 0; P.xchg.nranks = 1;
        if (kp.reduce_mode != RM_NONE) {
            rt.fill_exchange(P.xchg, &P.ticket);                                          // sharded run: ticket = the exchange's own sequence
            double* slot = rt.reduce_slot >= 0 ? rt.h_ticket_dev + 4 * rt.reduce_slot : nullptr;
            rt.last_tape_xhost = P.xchg.nranks > 1 && P.xchg.host_table != nullptr;
            if (rt.last_tape_xhost) { /* published in the shared host table */ }
            else if (P.xchg.nranks > 1) P.host_result = slot;
            else if (rt.comm_size == 1 && rt.opt.zero_copy_reduce && slot) { P.ticket = (rt.reduce_ticket += 1.0); P.host_result = slot; }
            rt.last_tape_ticket = (P.host_result || rt.last_tape_xhost) ? P.ticket : 0.0;
        }
        for (size_t k = 0; k < kp.ptr_local.size(); k++) P.ptrs[k] = info[kp.ptr_local[k]].buf;
        std::memcpy(P.instr, kp.words.data(), sizeof(TapeInstr) * kp.words.size());
        if (patch)
            for (const ImmPatch& ip : kp.patches) {
                const uint32_t bits = f2u(rt.nodes[info[ip.src >> 2].node].imm[ip.src & 3]);
                P.instr[ip.word].y = ip.neg ? (bits ^ 0x80000000u) : bits;
            }
        if (rt.opt.profile) rt.profile_begin();
        const auto t_launch0 = std::chrono::steady_clock::now();
        FMC_CUDA(launch_tape(P, kp.grid, kp.n_warps, rt.stream, rt.opt.tape_upload_stream ? rt.copy_stream : nullptr));
        rt.hostprof.launch += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_launch0).count();
        if (rt.opt.profile) rt.profile_end(4ull * (uint64_t)n * (uint64_t)(kp.n_leaf_slots + kp.n_result_stores), 4ull * (uint64_t)n * (uint64_t)kp.ptr_local.size());
        rt.stats.n_kernels++; rt.stats.n_tape_kernels++; rt.stats.n_tape_instr += (uint64_t)kp.n_instr;
    }
};

}  // namespace

void tape_cache_clear() { g_cache.clear(); }
void tape_cache_stats(uint64_t* hits, uint64_t* misses, uint64_t* entries) { *hits = g_cache.hits; *misses = g_cache.misses; *entries = g_cache.entries; }

// Windows (option window_levels). A flush that covers several dependency levels of still-referenced values — the time steps of a
// simulation, each reading what the step before stored — is cut into windows of W levels, and inside a window the targets of the
// LAST level are visited first. The post-order walk of run_cone then emits every chain (a model component) through all W levels
// before it starts the next chain, so what a level hands to the next one stays in the accumulator / register file and only the
// state shared between chains (running sums, one per level) waits in register-file slots: a value a step stores is not read back
// from HBM by the next step of the same window. Level of a target = the longest chain of still-referenced pending values below it.
// hold_last (automatic flushes): the flush threshold falls in the middle of a time step and of a window; what is incomplete stays
// pending, so that windows begin and end on whole levels. Returns true when something was held back.
bool Runtime::run_windows(const std::vector<int32_t>& targets, const std::vector<int32_t>& recorded, bool hold_last) {
    // W adapts: a model whose chains share more state per level than the register file holds (a three-factor LMM: three running sums
    // and three Brownian increments per time step) makes the generator spill and cut a window into many small launches — more than
    // two per level, i.e. worse than no window. Such a window lowers W for the flushes that follow (never below 1: one launch per
    // level, the order of round 1); setting the option window_levels starts over.
    if (window_levels_now <= 0 || window_levels_now > opt.window_levels) window_levels_now = opt.window_levels;
    const int W = std::max(1, window_levels_now);
    epoch++;
    if (epoch == 0) { for (auto& nd : nodes) nd.epoch = 0; epoch = 1; }
    // levels in ONE linear pass over the pending list: it is in recording order, so a node's operands come before it (an entry whose
    // slot was recycled by a later node can come before ITS operands: such a node is walked depth-first on the spot)
    static thread_local std::vector<std::pair<int32_t, int>> lstack;
    int max_level = 0;
    auto level_of = [&](const Node& nd) {
        int32_t lev = 0;
        for (int k = 0; k < 3; k++) {
            const int32_t u = nd.in[k];
            if (u >= 0 && nodes[u].state == NS_LAZY) lev = std::max(lev, nodes[u].local + (nodes[u].ext_refs > 0 ? 1 : 0));
        }
        return lev;
    };
    for (int32_t t : recorded) {
        Node& nt = nodes[t];
        if (nt.state != NS_LAZY || nt.epoch == epoch) continue;
        bool ready = true;
        for (int k = 0; k < 3; k++) { const int32_t u = nt.in[k]; if (u >= 0 && nodes[u].state == NS_LAZY && nodes[u].epoch != epoch) ready = false; }
        if (ready) {
            nt.epoch = epoch; nt.local = level_of(nt);
            max_level = std::max(max_level, (int)nt.local);
            continue;
        }
        nt.epoch = epoch; nt.local = 0;
        lstack.clear();
        lstack.emplace_back(t, 0);
        while (!lstack.empty()) {
            auto& top = lstack.back();
            const int32_t v = top.first;
            if (top.second < 3) {
                const int32_t u = nodes[v].in[top.second++];
                if (u >= 0 && nodes[u].state == NS_LAZY && nodes[u].epoch != epoch) { nodes[u].epoch = epoch; nodes[u].local = 0; lstack.emplace_back(u, 0); }
            } else {
                nodes[v].local = level_of(nodes[v]);
                max_level = std::max(max_level, (int)nodes[v].local);
                lstack.pop_back();
            }
        }
    }
    // Windows pay when a level is WIDE — many chains advancing side by side, each handing its value to the next level — so that one
    // launch per window replaces one per level. A deep, narrow graph (one chain with every intermediate still referenced: a
    // Black-Scholes path, the forward pass of an AAD tape) would be cut into a launch per W levels instead of one per flush:
    // below eight still-referenced values per level the flush runs as one cone, as it always did.
    // Nor do they pay when most pending values are still referenced (a differentiable wrapper keeps every intermediate for its
    // reverse sweep: a "level" is then one operation, not a time step, and everything is stored anyway; measured: the forward pass
    // of an AAD LMM vega ran 2x slower in windows).
    static thread_local std::vector<char> level_used;
    level_used.assign((size_t)max_level + 1, 0);
    long long n_levels_used = 0, n_targets = 0;
    // (the pending list names a recycled node slot once per life: count each target once)
    epoch++;
    if (epoch == 0) { for (auto& nd : nodes) nd.epoch = 0; epoch = 1; }
    for (int32_t t : targets) {
        Node& nt = nodes[t];
        if (nt.state != NS_LAZY || nt.epoch == epoch) continue;
        nt.epoch = epoch;
        n_targets++;
        if (!level_used[(size_t)nt.local]) { level_used[(size_t)nt.local] = 1; n_levels_used++; }
    }
    if (n_targets < 8 * n_levels_used || 4 * n_targets > (long long)n_lazy) {
        static const bool log_cone = std::getenv("FMC_LOG_TAPES") != nullptr;
        if (log_cone) std::fprintf(stderr, "[fmc windows] one cone: %lld targets, %lld levels used of %d, %lld pending\n", n_targets, n_levels_used, max_level + 1, (long long)n_lazy);
        run_cone(targets, nullptr);
        return false;
    }
    const int n_win = max_level / W + 1;
    // held back: the top three levels and with them the window they belong to. The handles a caller holds only for the duration of
    // a step (its vector of drifts, the running state of the component being updated) are still-referenced values too: they make
    // the step in progress look up to three levels deep, and flushing them would store and re-read values nobody asks for
    const int n_run = hold_last ? std::max(0, (max_level - 2) / W) : n_win;
    if (n_run == 0) return true;
    if (n_win == 1) {
        run_cone(targets, nullptr);
        return false;
    }
    std::vector<std::vector<int32_t>> win((size_t)n_win);
    // visiting order inside a window: first the ends of the chains — targets that no pending value reads (a component that leaves
    // the simulation at a lower level) and the targets of the window's last level — then the rest, each in recording order
    for (int pass = 0; pass < 2; pass++)
        for (int32_t t : targets) {
            if (nodes[t].state != NS_LAZY) continue;
            const int lev = nodes[t].local, w = lev / W;
            const bool end = lev == std::min(max_level, w * W + W - 1) || nodes[t].int_refs == 0;
            if (end == (pass == 0)) win[(size_t)w].push_back(t);
        }
    static const bool log_windows = std::getenv("FMC_LOG_TAPES") != nullptr;
    if (log_windows) {
        std::fprintf(stderr, "[fmc windows] %zu targets, %lld pending, levels 0..%d, W=%d, runs %d of:", targets.size(), (long long)n_lazy, max_level, W, n_run);
        for (auto& wt : win) std::fprintf(stderr, " %zu", wt.size());
        std::vector<int> per_level((size_t)max_level + 1, 0);
        for (int32_t t : targets) if (nodes[t].state == NS_LAZY) per_level[(size_t)nodes[t].local]++;
        std::fprintf(stderr, "   per level:");
        for (int c : per_level) std::fprintf(stderr, " %d", c);
        std::fprintf(stderr, "\n");
    }
    struct Guard { bool& f; explicit Guard(bool& f_) : f(f_) { f = true; } ~Guard() { f = false; } } guard(windowing);
    struct Clear { std::unordered_set<const float*>& s; ~Clear() { s.clear(); } } clear_stored{window_stored};
    for (int w = 0; w < n_run; w++) {
        if (win[(size_t)w].empty()) continue;
        const uint64_t launches0 = stats.n_tape_kernels;
        run_cone(win[(size_t)w], nullptr);
        const int levels_here = std::min(W, max_level + 1 - w * W);
        // (a launch or two more than levels is normal: the chains that end inside the window, a ragged first level)
        if (W > 1 && stats.n_tape_kernels - launches0 > 2u * (uint64_t)levels_here + 2u && window_levels_now == W) {
            window_levels_now = W - 1;
            if (log_windows) std::fprintf(stderr, "[fmc windows] a window of %d levels took %llu launches: W -> %d\n", levels_here,
                                          (unsigned long long)(stats.n_tape_kernels - launches0), window_levels_now);
        }
        if (w + 1 < n_run) for (int32_t t : win[(size_t)w]) if (nodes[t].state == NS_MAT) window_stored.insert(nodes[t].buf);
    }
    return n_run < n_win;
}

void Runtime::run_cone(const std::vector<int32_t>& targets, const ReduceSpec* red) {
    require_init();
    struct Timer {
        HostProfile& hp; double launch0; std::chrono::steady_clock::time_point t0;
        explicit Timer(HostProfile& h) : hp(h), launch0(h.launch), t0(std::chrono::steady_clock::now()) {}
        ~Timer() { hp.codegen += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() - (hp.launch - launch0); }
    } timer(hostprof);
    epoch++;
    if (epoch == 0) { for (auto& nd : nodes) nd.epoch = 0; epoch = 1; }
    stats.n_flushes++;

    struct Phase {                       // FMC_HOST_PROFILE only: no allocation, no clock call otherwise
        int id = -1; std::chrono::steady_clock::time_point t0;
        void reset(int next) {
            if (!g_phase.on) return;
            const auto now = std::chrono::steady_clock::now();
            if (id >= 0) g_phase.us[id] += std::chrono::duration<double, std::micro>(now - t0).count();
            id = next; t0 = now;
        }
        ~Phase() { reset(-1); }
    } ph;
    ph.reset(0);
    // ---- 1. collect the cone of lazy nodes, in depth-first post-order from the targets ----
    // Post-order (operands first, each value as late as its first consumer allows) keeps few intermediates alive at a
    // time: e.g. an Euler step whose caller computed all 80 drifts before applying any of them is emitted component
    // by component, so the register file holds a handful of values instead of 80.
    // One pass over the nodes does everything that needs them: when a node is finished (all operands numbered) it gets
    // its cone position, its operands' use counts go up, materialised operands are numbered in order of first use, and
    // the node's words of the tape-cache key are written (see 1b).
    int64_t n = -1;
    for (int32_t t : targets) {
        const Node& nd = nodes[t];
        if (n < 0) n = nd.n;
        else if (nd.n != n) {
            // targets of different sizes: run them as separate cones
            std::vector<int32_t> same, other;
            for (int32_t u : targets) (nodes[u].n == n ? same : other).push_back(u);
            run_cone(same, red);
            run_cone(other, red);
            return;
        }
    }
    Gen g(*this, n);
    // the per-cone vectors keep their capacity from flush to flush (a calibration runs thousands of cones per second: growing
    // them from empty every time was a sixth of the host time of a step)
    static thread_local std::vector<Info> info_keep;
    static thread_local std::vector<int32_t> use_keep;
    struct Keep {
        Gen& g; std::vector<Info>& i; std::vector<int32_t>& u;
        Keep(Gen& g_, std::vector<Info>& i_, std::vector<int32_t>& u_) : g(g_), i(i_), u(u_) { i.clear(); u.clear(); g.info.swap(i); g.use_list.swap(u); }
        ~Keep() { g.info.swap(i); g.use_list.swap(u); }
    } keep(g, info_keep, use_keep);
    constexpr int32_t LEAF_BASE = 0x40000000;            // Node::local of a leaf until the cone size is known: LEAF_BASE + ordinal
    static thread_local std::vector<std::pair<int32_t, int>> stack;
    static thread_local std::vector<int32_t> leaf_nodes, leaf_uses;
    static thread_local std::vector<uint32_t> key;
    static thread_local std::vector<uint32_t> int_refs;  // per cone position
    stack.clear(); leaf_nodes.clear(); leaf_uses.clear(); key.clear(); int_refs.clear();
    key.resize(6);
    auto leaf_ordinal = [&](int32_t u) -> int32_t {
        Node& nd = nodes[u];
        if (nd.epoch != epoch) {
            nd.epoch = epoch;
            nd.local = LEAF_BASE + (int32_t)leaf_nodes.size();
            leaf_nodes.push_back(u); leaf_uses.push_back(0);
        }
        return nd.local - LEAF_BASE;
    };
    auto finish = [&](int32_t v) {
        Node& nd = nodes[v];
        const int32_t L = (int32_t)g.info.size();
        nd.local = L;
        Info f; f.node = v; f.lazy = true;
        g.info.push_back(f);
        int_refs.push_back(nd.int_refs);
        uint32_t w = (uint32_t)nd.op;
        if (nd.ext_refs > 0) w |= 1u << 8;
        uint32_t o[3];
        for (int k = 0; k < 3; k++) {
            const int32_t u = nd.in[k];
            if (u < 0) { o[k] = 0xffffffffu; if (nd.imm[k] == 1.0f) w |= 1u << (11 + k); continue; }
            if (nodes[u].state == NS_LAZY) { const int32_t lu = nodes[u].local; g.info[lu].uses++; o[k] = (uint32_t)lu; }
            else { const int32_t j = leaf_ordinal(u); leaf_uses[(size_t)j]++; o[k] = 0x80000000u | (uint32_t)j; }
        }
        key.push_back(w); key.push_back(o[0]); key.push_back(o[1]); key.push_back(o[2]);
    };
    for (int32_t t : targets) {
        if (nodes[t].state != NS_LAZY || nodes[t].epoch == epoch) continue;
        nodes[t].epoch = epoch;
        stack.emplace_back(t, 0);
        while (!stack.empty()) {
            auto& top = stack.back();
            const int32_t v = top.first;
            if (top.second < 3) {
                const int32_t u = nodes[v].in[top.second++];
                if (u >= 0 && nodes[u].state == NS_LAZY && nodes[u].epoch != epoch) { nodes[u].epoch = epoch; stack.emplace_back(u, 0); }
            } else {
                stack.pop_back();
                finish(v);
            }
        }
    }
    const int32_t n_cone = (int32_t)g.info.size();
    const int32_t T = red ? targets[0] : -1;
    int32_t weight_local = -1, target_local = -1;
    if (red && red->weight >= 0) { const int32_t j = leaf_ordinal(red->weight); leaf_uses[(size_t)j]++; weight_local = n_cone + j; }
    if (red) {
        if (nodes[T].state == NS_LAZY) { target_local = nodes[T].local; g.info[target_local].uses++; }   // the reduction epilogue reads it from acc
        else { const int32_t j = leaf_ordinal(T); leaf_uses[(size_t)j]++; target_local = n_cone + j; }
    }
    for (size_t j = 0; j < leaf_nodes.size(); j++) {
        Node& nd = nodes[leaf_nodes[j]];
        nd.local = n_cone + (int32_t)j;
        Info f; f.node = leaf_nodes[j]; f.lazy = false; f.buf = nd.buf; f.uses = leaf_uses[j];
        g.info.push_back(f);
    }

    // ---- 1b. tape cache: a cone with the same structure was lowered before -> replay its launches with this cone's
    // buffers and immediates. The key holds everything the code generator looks at except the immediates' values (and
    // whether one equals 1.0f, which the ACCRUE / DISCOUNT fusion tests): per node the operation, where each operand
    // comes from (scalar / cone position / number of the leaf in order of first use), whether the caller or a pending
    // node outside the cone still refers to it, whether it is a target; the vector length, the reduction. Options that
    // steer the generator empty the cache when they change (capi.cpp).
    ConePlan* hit = nullptr;
    std::unique_ptr<ConePlan> fresh;
    uint64_t hash = 0;
    if (opt.tape_cache && n_cone > 0) {
        key[0] = (uint32_t)n; key[1] = (uint32_t)((uint64_t)n >> 32);
        key[2] = red ? (uint32_t)red->mode : 0u;
        key[3] = (uint32_t)target_local; key[4] = (uint32_t)weight_local;
        key[5] = (uint32_t)n_cone;
        for (int32_t L = 0; L < n_cone; L++) {
            const int32_t cone_uses = g.info[L].uses - ((red && L == target_local) ? 1 : 0);
            if ((int64_t)int_refs[(size_t)L] > (int64_t)cone_uses) key[6 + 4 * (size_t)L] |= 1u << 9;
        }
        for (int32_t t : targets) if (nodes[t].state == NS_LAZY) key[6 + 4 * (size_t)nodes[t].local] |= 1u << 10;
        hash = 1469598103934665603ull;
        for (uint32_t w : key) { hash ^= w; hash *= 1099511628211ull; }
        auto it = g_cache.map.find(hash);
        if (it != g_cache.map.end())
            for (auto& cp : it->second) if (cp->key == key) { hit = cp.get(); break; }
        if (hit) g_cache.hits++;
        else { g_cache.misses++; fresh.reset(new ConePlan); fresh->key = key; g.keep_plans = true; }
    }
    if (hit) {
        ph.reset(5);
        // buffers of the nodes this cone stores (or spills), all up front
        std::vector<float*> got;
        try {
            for (int32_t L = 0; L < n_cone; L++)
                if (hit->node_flags[(size_t)L] & 1u) {
                    g.info[L].buf = (float*)pool.alloc(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
                    got.push_back(g.info[L].buf);
                }
        } catch (...) {
            for (float* b : got) pool.free(b);
            throw;
        }
        for (int32_t L = 0; L < n_cone; L++) g.info[L].eph = (hit->node_flags[(size_t)L] & 2u) != 0;
        for (const KernelPlan& kp : hit->kernels) g.submit(kp, red ? red->param : 0.0, true);
    } else {
    ph.reset(1);
    // ---- 2. store / ephemeral classification (reverse topological order) ----
    for (int32_t L = n_cone - 1; L >= 0; L--) {
        Info& f = g.info[L];
        const Node& nd = nodes[f.node];
        int32_t cone_uses = f.uses - ((red && L == target_local) ? 1 : 0);
        const bool outside = (int64_t)nd.int_refs > (int64_t)cone_uses;
        if (red && L == target_local) f.eph = !outside;
        else f.eph = red && nd.ext_refs == 0 && !outside && cone_uses > 0 && f.eph_uses == cone_uses;
        if (f.eph) {
            for (int k = 0; k < 3; k++) {
                const int32_t u = nd.in[k];
                if (u >= 0 && nodes[u].state == NS_LAZY) g.info[nodes[u].local].eph_uses++;
            }
            f.store = false;
        } else {
            // written to HBM iff somebody can still ask for it after this flush
            f.store = nd.ext_refs > 0 || outside || f.eph_uses > 0;
        }
    }
    if (!red) for (int32_t t : targets) if (nodes[t].state == NS_LAZY) g.info[nodes[t].local].store = true;

    // use lists: for every value the cone positions of the nodes that read it, one entry per operand slot
    // (the reduction epilogue counts as position n_cone)
    {
        int32_t off = 0;
        for (Info& f : g.info) { f.ubeg = f.ucur = off; off += f.uses; f.uend = f.ubeg; }
        g.use_list.assign((size_t)off, 0);
        auto add_use = [&](int32_t lu, int32_t at) { Info& f = g.info[lu]; g.use_list[(size_t)f.uend++] = at; };
        for (int32_t L = 0; L < n_cone; L++) {
            const Node& nd = nodes[g.info[L].node];
            for (int k = 0; k < 3; k++) if (nd.in[k] >= 0) add_use(nodes[nd.in[k]].local, L);
        }
        if (weight_local >= 0) add_use(weight_local, n_cone);
        if (target_local >= 0) add_use(target_local, n_cone);
    }

    ph.reset(2);
    // ---- 3. emit ----
    g.begin_kernel();
    for (int32_t L = 0; L < n_cone; L++) {
        // margins: one node emits < 16 instructions / < 8 new pointers; a cut stores at most TAPE_REGS + 1 live values;
        // ring scheduling adds at most two instructions per leaf reference plus two per ring slot
        // (outside a window the kernels stay at the size the default path was tuned for)
        const int instr_limit = windowing ? TAPE_MAX_INSTR : std::min(TAPE_MAX_INSTR, 2046), ptr_limit = windowing ? TAPE_MAX_PTRS : std::min(TAPE_MAX_PTRS, 384);
        if ((int)g.A.size() + 2 * g.n_leaf_refs + 2 * TAPE_MAX_RING + 64 > instr_limit || (int)g.ptrs.size() + 28 > ptr_limit) g.cut();
        else if (!g.operands_available(L)) g.cut();
        g.emit_node(L);
    }
    if (red) {
        g.pos = n_cone;
        if (!g.available(target_local)) g.cut();
        g.set_pins(target_local, weight_local);
        if (g.acc_owner != target_local) g.take_acc(target_local, 1);
        g.launch(red->mode, red->param, weight_local);
    } else {
        g.launch(RM_NONE, 0.0, -1);
    }

    if (fresh) {
        fresh->kernels = std::move(g.plans);
        fresh->node_flags.resize((size_t)n_cone);
        size_t words = 0;
        for (int32_t L = 0; L < n_cone; L++) fresh->node_flags[(size_t)L] = (uint8_t)((g.info[L].buf ? 1u : 0u) | (g.info[L].eph ? 2u : 0u));
        for (const KernelPlan& kp : fresh->kernels) words += kp.words.size();
        if (g_cache.entries >= 8192 || g_cache.words + words > (64u << 20) / sizeof(TapeInstr)) g_cache.clear();   // bounded: start over
        g_cache.entries++; g_cache.words += words;
        g_cache.map[hash].push_back(std::move(fresh));
    }
    }   // cache miss

    ph.reset(6);
    // ---- 4. bookkeeping: stored / spilled nodes become materialised and drop their operands ----
    for (int32_t L = 0; L < n_cone; L++) {
        Info& f = g.info[L];
        const int32_t v = f.node;
        Node& nd = nodes[v];
        if (nd.state != NS_LAZY) continue;          // already freed by a cascade
        if (f.buf) {
            nd.buf = f.buf; nd.state = NS_MAT; nd.op = N_LEAF;
            n_lazy--;
            stats.n_stored++;
            int32_t ins3[3] = {nd.in[0], nd.in[1], nd.in[2]};
            nd.in[0] = nd.in[1] = nd.in[2] = -1;
            for (int k = 0; k < 3; k++) if (ins3[k] >= 0) release_int(ins3[k]);
            if (nodes[v].ext_refs == 0 && nodes[v].int_refs == 0) maybe_free(v);
        } else if (!f.eph) {
            stats.n_fused++;
        }
    }
}

bool Runtime::reduce(int32_t idx, const ReduceSpec& spec_in, double out[3]) {
    require_init();
    ReduceSpec spec = spec_in;
    if (spec.weight >= 0) materialize(spec.weight);
    const bool empty = nodes[idx].n == 0;
    if (empty && comm_size == 1) { out[0] = 0.0; out[1] = NAN; out[2] = NAN; return false; }
    // batched averages (runtime.h): a sum that an earlier batch already delivered, or the sums of idx and of the vectors that
    // were materialised together with it, in one launch
    if (opt.batch_reduce && comm_size == 1 && spec.mode == RM_SUM && spec.weight < 0 && !empty) {
        auto pf = prefetched.find(idx);
        if (pf != prefetched.end() && pf->second.gen == nodes[idx].gen && nodes[idx].state == NS_MAT) {
            out[0] = (double)nodes[idx].n; out[1] = pf->second.sum; out[2] = 0.0;
            reduce_streak = true;
            return false;
        }
        if (nodes[idx].state == NS_LAZY && reduce_streak) flush_all();     // the tail of a series of averages: same treatment
        if (nodes[idx].state == NS_MAT && opt.leaf_reduce_kernel && reduce_batch(idx, out)) { reduce_streak = true; return false; }
    }
    const bool p2p = use_p2p();
    bool xhost_launch = false;                // this reduction's kernel publishes into the shared host table
    // the result slot of this reduction (released when the result has been read, also on the error paths)
    struct Slot {
        Runtime& rt; int s = -1;
        explicit Slot(Runtime& r) : rt(r) {
            if (~rt.ticket_slots_busy) { s = __builtin_ctzll(~rt.ticket_slots_busy); rt.ticket_slots_busy |= 1ull << s; }
            rt.reduce_slot = s;
        }
        ~Slot() { if (s >= 0) rt.ticket_slots_busy &= ~(1ull << s); }
    } slot(*this);
    double ticket = 0.0;                      // what the host spins on (0: result comes by copy + stream synchronisation)
    if (empty || (nodes[idx].state == NS_MAT && opt.leaf_reduce_kernel)) {
        // nothing to interpret: plain streaming reduction (reduce_kernel.cu); an empty slice of a sharded vector
        // contributes {0, 0, 0} and still takes part in the exchange
        ReduceParams P;
        P.n = nodes[idx].n; P.mode = spec.mode; P.param = spec.param;
        P.x = empty ? nullptr : nodes[idx].buf; P.w = (!empty && spec.weight >= 0) ? nodes[spec.weight].buf : nullptr;
        P.partials = d_partials; P.counter = d_counter; P.result = d_result;
        P.host_result = nullptr; P.ticket = 0.0;
        fill_exchange(P.xchg, &P.ticket);
        double* slot = reduce_slot >= 0 ? h_ticket_dev + 4 * reduce_slot : nullptr;
        xhost_launch = P.xchg.nranks > 1 && P.xchg.host_table != nullptr;
        if (xhost_launch) { /* published in the shared host table */ }
        else if (P.xchg.nranks > 1) P.host_result = slot;
        else if (comm_size == 1 && opt.zero_copy_reduce && slot) { P.ticket = (reduce_ticket += 1.0); P.host_result = slot; }
        ticket = (P.host_result || xhost_launch) ? P.ticket : 0.0;
        const int64_t tiles = (P.n + reduce_tile_elems() - 1) / reduce_tile_elems();
        int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)sm_count * 8));
        grid = std::min(grid, max_grid);
        if (opt.grid_limit > 0) grid = std::min(grid, opt.grid_limit);
        if (opt.profile) profile_begin();
        const auto t_launch0 = std::chrono::steady_clock::now();
        FMC_CUDA(launch_reduce(P, grid, stream));
        hostprof.launch += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_launch0).count();
        if (opt.profile) profile_end(4ull * (uint64_t)P.n * (P.w ? 2u : 1u));
        stats.n_kernels++; stats.n_flushes++;
    } else {
        // The pending part of a simulation that this valuation reads (what the last automatic flush held back, and the steps
        // recorded since) goes through the windows first: the still-referenced pending values below idx, in recording order.
        // The valuation's own chain stays fused with its reduction; pending values it does not read stay pending.
        if (opt.fuse && opt.window_levels > 0 && n_lazy > opt.window_reduce_min && !windowing) {
            epoch++;
            if (epoch == 0) { for (auto& nd : nodes) nd.epoch = 0; epoch = 1; }
            static thread_local std::vector<int32_t> walk, below;
            walk.clear(); below.clear();
            nodes[idx].epoch = epoch; walk.push_back(idx);
            size_t n_ext = 0;
            while (!walk.empty()) {
                const int32_t v = walk.back(); walk.pop_back();
                for (int k = 0; k < 3; k++) {
                    const int32_t u = nodes[v].in[k];
                    if (u < 0 || nodes[u].state != NS_LAZY || nodes[u].epoch == epoch) continue;
                    nodes[u].epoch = epoch; walk.push_back(u);
                    if (nodes[u].ext_refs > 0) n_ext++;
                }
            }
            if (n_ext >= 32) {                                  // a simulation's worth, not the few cached numeraires a valuation extends
                const uint32_t mark = epoch;
                for (int32_t t : pending) if (t != idx && nodes[t].state == NS_LAZY && nodes[t].epoch == mark && nodes[t].ext_refs > 0) {
                    below.push_back(t);
                    nodes[t].epoch = mark - 1;                 // a recycled slot can appear twice in the list
                }
                static thread_local std::vector<int32_t> order;
                order = pending;                               // run_cone compacts nothing, but the walk must not alias a list that changes
                if (below.size() >= 32) run_windows(below, order, false);
            }
        }
        std::vector<int32_t> t{idx};
        last_tape_xhost = false;
        run_cone(t, &spec);                   // the fused chain -> reduce launch sets params->ticket (Gen::launch)
        ticket = last_tape_ticket;
        xhost_launch = last_tape_xhost;
    }
    const auto t_sync0 = std::chrono::steady_clock::now();
    const uint64_t stamp_at_launch = pool.free_stamp;   // blocks freed so far were last used by work queued before this kernel
    if (xhost_launch) {
        // Every rank's kernel stores its {count, value, M2} and then the ticket into slot [ticket % XSLOTS][rank] of the table in
        // shared host memory; every rank's host waits here for the R tickets and hands the partials to the caller's merge in rank
        // order (capi.cpp: merge_ranks) — the same arithmetic on every rank. No kernel waits for a peer and the compute stream is
        // free for whatever was queued behind the reduction. A rank cannot run ahead of the others by more than one reduction, so
        // a slot is not reused before everybody has read it.
        const int xs = (int)((long long)ticket % XSLOTS);
        const auto t_dead = t_sync0 + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(opt.exchange_timeout_s));
        for (int r = 0; r < comm_size; r++) {
            volatile double* t = xhost + ((size_t)xs * XMAX_RANKS + (size_t)r) * 4;
            unsigned spins = 0;
            while (t[3] != ticket) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
                if ((++spins & 0xfffu) == 0u) {
                    if (r == comm_rank) {                  // my own kernel: has it failed?
                        const cudaError_t q = cudaStreamQuery(stream);
                        if (q != cudaSuccess && q != cudaErrorNotReady) FMC_CUDA(q);
                    }
                    if (std::chrono::steady_clock::now() > t_dead)
                        fail(FMC_ERR_COMM, "reduction exchange timed out after %.0f s: rank %d did not deliver its partial (option exchange_timeout_s)", opt.exchange_timeout_s, r);
                }
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            h_result[4 * r] = t[0]; h_result[4 * r + 1] = t[1]; h_result[4 * r + 2] = t[2];
        }
        out[0] = h_result[4 * comm_rank]; out[1] = h_result[4 * comm_rank + 1]; out[2] = h_result[4 * comm_rank + 2];
        settled_stamp = std::max(settled_stamp, stamp_at_launch);
        hostprof.sync += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_sync0).count();
        stats.d2h += 32 * (uint64_t)comm_size;
        return false;
    }
    if (ticket != 0.0) {
        // the last block of the reduction wrote {count, value, M2} (after the in-kernel exchange: of ALL ranks) and then the
        // ticket into mapped pinned memory: spin on the ticket instead of a 32-byte copy plus a stream synchronisation
        // Single-rank runs wait WITHOUT the runtime lock: other host threads keep recording and launching meanwhile (their
        // reductions use other ticket slots), the way the reference's valuation threads share one device. Sharded runs keep
        // the lock: every rank must issue its reductions in the same order.
        volatile double* h = h_ticket + 4 * slot.s;
        RuntimeLock* mine = (comm_size == 1) ? held : nullptr;
        cudaStream_t s = stream;
        if (mine) { held = nullptr; mine->unlock(); }
        unsigned spins = 0;
        auto t_query = t_sync0;
        const char* err = nullptr; int err_code = 0; cudaError_t cuda_err = cudaSuccess;
        while (h[3] != ticket) {
            if (h[3] == -ticket) { err = "reduction exchange timed out: a peer rank did not deliver its partial"; err_code = FMC_ERR_COMM; break; }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();                  // be polite to the sibling hyper-thread (another valuation thread may be recording)
#endif
            if ((++spins & 0x3ffu) == 0u) {
                // liveness check, at most every 200 us: the driver call contends with the threads that are launching
                const auto now = std::chrono::steady_clock::now();
                if (now - t_query < std::chrono::microseconds(200)) continue;
                t_query = now;
                const cudaError_t q = cudaStreamQuery(s);
                if (q == cudaSuccess) { if (h[3] == ticket) break; err = "reduction finished without publishing its result"; err_code = FMC_ERR_CUDA; break; }
                if (q != cudaErrorNotReady) { cuda_err = q; break; }
            }
        }
        // sums of an unsharded vector come as {h[2] = value, h[3] = ticket} in one store (reduce_common.cuh: finish_reduction)
        const bool sum_pair = comm_size == 1 && (spec.mode == RM_SUM || spec.mode == RM_DOT || spec.mode == RM_WSQ);
        std::atomic_thread_fence(std::memory_order_acquire);
        const double r0 = sum_pair ? (double)nodes[idx].n : h[0], r1 = sum_pair ? h[2] : h[1], r2 = sum_pair ? 0.0 : h[2];
        if (mine) { mine->lock(); held = mine; }
        if (err) fail(err_code, "%s", err);
        FMC_CUDA(cuda_err);
        out[0] = r0; out[1] = r1; out[2] = r2;
        settled_stamp = std::max(settled_stamp, stamp_at_launch);
        hostprof.sync += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_sync0).count();
        stats.d2h += 32;
        return p2p;
    }
    if (comm_size > 1) {
        // NCCL exchange (peer tables unavailable): every rank's {count, value, M2} into one table, merged on the host in
        // rank order (capi.cpp)
        allgather(d_result, d_result + 8, 4);
        FMC_CUDA(cudaMemcpyAsync(h_result, d_result + 8, sizeof(double) * 4 * (size_t)comm_size, cudaMemcpyDeviceToHost, stream));
        sync_stream();
        hostprof.sync += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_sync0).count();
        stats.d2h += 32 * (uint64_t)comm_size;
        out[0] = h_result[4 * comm_rank]; out[1] = h_result[4 * comm_rank + 1]; out[2] = h_result[4 * comm_rank + 2];
        return false;
    }
    FMC_CUDA(cudaMemcpyAsync(h_result, d_result, sizeof(double) * 4, cudaMemcpyDeviceToHost, stream));
    sync_stream();
    hostprof.sync += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_sync0).count();
    stats.d2h += 32;
    out[0] = h_result[0]; out[1] = h_result[1]; out[2] = h_result[2];
    return false;
}

}  // namespace fmc
