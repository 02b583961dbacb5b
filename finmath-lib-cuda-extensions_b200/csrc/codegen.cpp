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
#include <cstring>

#include "runtime.h"

namespace fmc {

namespace {

struct Info {
    int32_t node;
    int32_t rem = 0;        // remaining uses (operand slots) inside this flush
    int32_t uses = 0;       // total uses inside the cone
    int32_t eph_uses = 0;
    int16_t slot = -1;      // pointer-table slot in the current kernel
    int8_t reg = -1;
    bool lazy = false;      // cone node (to be computed) vs. materialised leaf
    bool store = false, eph = false, computed = false;
    float* buf = nullptr;   // device buffer (existing for leaves, new for stored / spilled nodes)
    uint32_t last_use = 0;
};

uint32_t rev_op(uint32_t op) {
    switch (op) {
    case T_SUB: return T_BUS;
    case T_BUS: return T_SUB;
    case T_DIV: return T_VID;
    case T_VID: return T_DIV;
    default: return op;   // ADD MUL MIN MAX commute
    }
}

struct Gen {
    Runtime& rt;
    int64_t n;
    std::vector<Info> info;
    std::vector<TapeInstr> ins;
    std::vector<float*> ptrs;
    std::vector<int32_t> slotted;       // locals that own a slot in the current kernel
    int32_t reg_owner[TAPE_REGS];
    int32_t acc_owner = -1;
    int regs_used = 0;
    int n_leaf_slots = 0, n_result_stores = 0;   // algorithmic traffic of the current kernel (spills excluded)
    uint32_t tick = 0;
    TapeParams* params;                 // reused launch parameter block

    Gen(Runtime& r, int64_t n_) : rt(r), n(n_) {
        for (int j = 0; j < TAPE_REGS; j++) reg_owner[j] = -1;
        static thread_local TapeParams tp;
        params = &tp;
    }

    bool has_buf(int32_t L) const { return info[L].buf != nullptr; }

    void begin_kernel() {
        ins.clear(); ptrs.clear();
        for (int32_t L : slotted) info[L].slot = -1;
        slotted.clear();
        for (int j = 0; j < TAPE_REGS; j++) {
            if (reg_owner[j] >= 0) info[reg_owner[j]].reg = -1;
            reg_owner[j] = -1;
        }
        acc_owner = -1;
        regs_used = 0;
        n_leaf_slots = 0; n_result_stores = 0;
    }

    int slot_for(int32_t L) {
        Info& f = info[L];
        if (f.slot < 0) {
            f.slot = (int16_t)ptrs.size();
            ptrs.push_back(f.buf);
            slotted.push_back(L);
            if (!f.lazy) n_leaf_slots++;
        }
        return f.slot;
    }

    void ensure_buffer(int32_t L) {
        Info& f = info[L];
        if (!f.buf) f.buf = (float*)rt.pool.alloc(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
    }

    void emit(uint32_t op, uint32_t src, uint32_t idx, float imm = 0.f) { ins.push_back(enc(op, src, idx, imm)); }

    // write local L (currently in acc or in a register) to its HBM buffer
    void store_value(int32_t L) {
        ensure_buffer(L);
        const int slot = slot_for(L);
        uint32_t src;
        if (acc_owner == L) src = S_ACC;
        else src = S_REG0 + (uint32_t)info[L].reg;
        ins.push_back(enc_stg((uint32_t)slot, src));
    }

    void free_reg_of(int32_t L) {
        Info& f = info[L];
        if (f.reg >= 0) { reg_owner[f.reg] = -1; f.reg = -1; }
    }

    // find a register; may evict (and, if needed, spill to HBM) a value that is not pinned
    int32_t pins[3] = {-1, -1, -1};      // operands of the instruction being built: never evicted
    void set_pins(int32_t a, int32_t b = -1, int32_t c = -1) { pins[0] = a; pins[1] = b; pins[2] = c; }

    int alloc_reg() {
        for (int j = 0; j < TAPE_REGS; j++) if (reg_owner[j] < 0) { regs_used = std::max(regs_used, j + 1); return j; }
        int victim = -1;
        // prefer a value that already has an HBM copy, then the least recently used
        for (int pass = 0; pass < 2 && victim < 0; pass++) {
            uint32_t best = 0xffffffffu;
            for (int j = 0; j < TAPE_REGS; j++) {
                const int32_t L = reg_owner[j];
                if (L == pins[0] || L == pins[1] || L == pins[2]) continue;
                if (pass == 0 && !has_buf(L)) continue;
                if (info[L].last_use < best) { best = info[L].last_use; victim = j; }
            }
        }
        if (victim < 0) fail(FMC_ERR_UNSUPPORTED, "internal: register allocation failed");
        const int32_t L = reg_owner[victim];
        if (!has_buf(L)) store_value(L);        // spill
        free_reg_of(L);
        return victim;
    }

    void put_acc_in_reg() {
        const int32_t L = acc_owner;
        const int j = alloc_reg();
        emit(T_STR, 0, (uint32_t)j);
        reg_owner[j] = L; info[L].reg = (int8_t)j;
    }

    // the accumulator is about to be overwritten: keep its value if somebody still needs it
    void save_acc() {
        const int32_t L = acc_owner;
        if (L < 0) return;
        const Info& f = info[L];
        if (f.rem > 0 && f.reg < 0) {
            if (!has_buf(L)) put_acc_in_reg();
            else {
                // value is already in HBM: keep a register copy only if one is free
                for (int j = 0; j < TAPE_REGS; j++) if (reg_owner[j] < 0) { put_acc_in_reg(); break; }
            }
        }
    }

    void src_of(int32_t L, uint32_t& src, uint32_t& idx) {
        const Info& f = info[L];
        idx = 0;
        if (acc_owner == L) { src = S_ACC; return; }
        if (f.reg >= 0) { src = S_REG0 + (uint32_t)f.reg; return; }
        if (!f.buf) fail(FMC_ERR_UNSUPPORTED, "internal: operand has no location");
        src = S_LEAF; idx = (uint32_t)slot_for(L);
    }

    // make acc hold L; `uses_now` operand slots of L are consumed by the instruction being built
    void take_acc(int32_t L, int uses_now) {
        if (acc_owner != L) {
            save_acc();
            uint32_t src, idx;
            src_of(L, src, idx);
            emit(T_MOV, src, idx);
            acc_owner = L;
        }
        Info& f = info[L];
        if (f.rem > uses_now && f.reg < 0 && !has_buf(L)) put_acc_in_reg();
    }

    void consume(int32_t L) {
        Info& f = info[L];
        f.rem--;
        f.last_use = ++tick;
        if (f.rem <= 0) free_reg_of(L);
    }

    void emit_binary(uint32_t op, int32_t A, float immA, int32_t B, float immB) {
        // choose the operand that sits in (or goes to) the accumulator
        int32_t X, O; float immO; uint32_t xop;
        if (A >= 0 && acc_owner == A) { X = A; O = B; immO = immB; xop = op; }
        else if (B >= 0 && acc_owner == B) { X = B; O = A; immO = immA; xop = rev_op(op); }
        else if (A >= 0) { X = A; O = B; immO = immB; xop = op; }
        else { X = B; O = A; immO = immA; xop = rev_op(op); }
        const int uses_now = (O == X) ? 2 : 1;
        set_pins(X, O);
        take_acc(X, uses_now);
        if (O >= 0) {
            uint32_t src, idx;
            src_of(O, src, idx);
            emit(xop, src, idx);
        } else {
            emit(xop, S_IMM, 0, immO);
        }
        consume(X);
        if (O >= 0) consume(O);
    }

    void emit_node(int32_t L) {
        const Node& nd = rt.nodes[info[L].node];
        auto loc = [&](int k) -> int32_t { return nd.in[k] >= 0 ? rt.nodes[nd.in[k]].local : -1; };
        switch (nd.op) {
        case N_ADD: emit_binary(T_ADD, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_SUB: emit_binary(T_SUB, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MUL: emit_binary(T_MUL, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_DIV: emit_binary(T_DIV, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MIN: emit_binary(T_MIN, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_MAX: emit_binary(T_MAX, loc(0), nd.imm[0], loc(1), nd.imm[1]); break;
        case N_SQRT: case N_EXP: case N_LOG: case N_SIN: case N_COS: case N_ABS: case N_INV: case N_ISNAN: case N_POW: {
            static const uint32_t map[] = {0, 0, 0, 0, 0, 0, 0, T_SQRT, T_EXP, T_LOG, T_SIN, T_COS, T_ABS, T_INV, T_ISNAN, T_POW};
            const int32_t A = loc(0);
            set_pins(A);
            take_acc(A, 1);
            emit(map[nd.op], 0, 0, nd.imm[1]);
            consume(A);
            break;
        }
        case N_CHOOSE: {
            const int32_t X = loc(0), A = loc(1), B = loc(2);
            set_pins(X, A, B);
            take_acc(X, 1);
            emit(T_SETP, 0, 0);
            consume(X);
            // acc := A
            if (A >= 0) {
                if (acc_owner != A) {
                    save_acc();
                    uint32_t src, idx; src_of(A, src, idx);
                    emit(T_MOV, src, idx);
                    acc_owner = A;
                }
                if (info[A].rem > 1 && info[A].reg < 0 && !has_buf(A)) put_acc_in_reg();
            } else {
                save_acc();
                emit(T_MOV, S_IMM, 0, nd.imm[1]);
                acc_owner = -1;
            }
            if (B >= 0) {
                uint32_t src, idx; src_of(B, src, idx);
                emit(T_SEL, src, idx);
            } else {
                emit(T_SEL, S_IMM, 0, nd.imm[2]);
            }
            if (A >= 0) consume(A);
            if (B >= 0) consume(B);
            break;
        }
        case N_CONST:
            set_pins(-1);
            save_acc();
            emit(T_MOV, S_IMM, 0, nd.imm[0]);
            break;
        default: fail(FMC_ERR_UNSUPPORTED, "internal: cannot emit node op %d", (int)nd.op);
        }
        acc_owner = L;
        info[L].computed = true;
        info[L].last_use = ++tick;
        if (info[L].store) { store_value(L); n_result_stores++; }
    }

    // end the current kernel early: everything live in acc / registers that is still needed goes to HBM
    void cut() {
        if (acc_owner >= 0 && info[acc_owner].rem > 0 && !has_buf(acc_owner)) store_value(acc_owner);
        for (int j = 0; j < TAPE_REGS; j++) {
            const int32_t L = reg_owner[j];
            if (L >= 0 && info[L].rem > 0 && !has_buf(L)) store_value(L);
        }
        launch(RM_NONE, 0.0, -1);
        begin_kernel();
    }

    void launch(int reduce_mode, double reduce_param, int32_t weight_local) {
        if (reduce_mode == RM_DOT || reduce_mode == RM_WSQ) {
            uint32_t src, idx; src_of(weight_local, src, idx);
            emit(T_END, src, idx);
        } else emit(T_END, S_IMM, 0);
        if (ins.size() <= 1 && reduce_mode == RM_NONE) return;   // nothing to do
        if ((int)ins.size() > TAPE_MAX_INSTR + 1 || (int)ptrs.size() > TAPE_MAX_PTRS)
            fail(FMC_ERR_UNSUPPORTED, "internal: tape overflow (%zu instr, %zu ptrs)", ins.size(), ptrs.size());
        TapeParams& P = *params;
        P.n = n;
        P.n_instr = (int)ins.size();
        P.reduce_mode = reduce_mode;
        P.reduce_param = reduce_param;
        P.partials = rt.d_partials;
        P.counter = rt.d_counter;
        P.result = rt.d_result;
        std::memcpy(P.ptrs, ptrs.data(), sizeof(float*) * ptrs.size());
        std::memcpy(P.instr, ins.data(), sizeof(TapeInstr) * ins.size());
        P.instr[ins.size()] = enc(T_END, S_IMM, 0, 0.f);       // the interpreter prefetches one word ahead
        const int64_t tiles = (n + TAPE_TILE - 1) / TAPE_TILE;
        static int blocks_fast = 0, blocks_smem[TAPE_REGS + 1] = {0};
        int per_sm;
        if (regs_used <= TAPE_REGS_FAST) { if (!blocks_fast) blocks_fast = tape_max_blocks_per_sm(regs_used); per_sm = blocks_fast; }
        else { if (!blocks_smem[regs_used]) blocks_smem[regs_used] = tape_max_blocks_per_sm(regs_used); per_sm = blocks_smem[regs_used]; }
        int grid = (int)std::min<int64_t>(tiles, (int64_t)per_sm * rt.sm_count);
        grid = std::min(grid, rt.max_grid);
        if (grid < 1) grid = 1;
        if (rt.opt.profile) rt.profile_begin();
        FMC_CUDA(launch_tape(P, grid, regs_used, rt.stream));
        if (rt.opt.profile) rt.profile_end(4ull * (uint64_t)n * (uint64_t)(n_leaf_slots + n_result_stores));
        rt.stats.n_kernels++; rt.stats.n_tape_kernels++; rt.stats.n_tape_instr += ins.size();
    }
};

}  // namespace

void Runtime::run_cone(const std::vector<int32_t>& targets, const ReduceSpec* red) {
    require_init();
    epoch++;
    if (epoch == 0) { for (auto& nd : nodes) nd.epoch = 0; epoch = 1; }
    stats.n_flushes++;

    // ---- 1. collect the cone of lazy nodes ----
    std::vector<int32_t> cone, stack;
    int64_t n = -1;
    for (int32_t t : targets) {
        Node& nd = nodes[t];
        if (n < 0) n = nd.n;
        else if (nd.n != n) {
            // targets of different sizes: run them as separate cones
            std::vector<int32_t> same, other;
            for (int32_t u : targets) (nodes[u].n == n ? same : other).push_back(u);
            run_cone(same, red);
            run_cone(other, red);
            return;
        }
        if (nd.state == NS_LAZY && nd.epoch != epoch) { nd.epoch = epoch; stack.push_back(t); }
    }
    while (!stack.empty()) {
        const int32_t v = stack.back(); stack.pop_back();
        cone.push_back(v);
        const Node& nd = nodes[v];
        for (int k = 0; k < 3; k++) {
            const int32_t u = nd.in[k];
            if (u >= 0 && nodes[u].state == NS_LAZY && nodes[u].epoch != epoch) { nodes[u].epoch = epoch; stack.push_back(u); }
        }
    }
    std::sort(cone.begin(), cone.end(), [&](int32_t a, int32_t b) { return nodes[a].seq < nodes[b].seq; });

    Gen g(*this, n);
    g.info.reserve(cone.size() * 2 + 4);
    for (int32_t v : cone) {
        nodes[v].local = (int32_t)g.info.size();
        Info f; f.node = v; f.lazy = true;
        g.info.push_back(f);
    }
    const int32_t n_cone = (int32_t)cone.size();
    auto leaf_local = [&](int32_t u) -> int32_t {
        Node& nd = nodes[u];
        if (nd.epoch != epoch || nd.local < 0 || nd.local >= (int32_t)g.info.size() || g.info[nd.local].node != u) {
            nd.epoch = epoch;
            nd.local = (int32_t)g.info.size();
            Info f; f.node = u; f.lazy = false; f.buf = nd.buf;
            g.info.push_back(f);
        }
        return nd.local;
    };
    // uses
    for (int32_t L = 0; L < n_cone; L++) {
        const Node& nd = nodes[cone[L]];
        for (int k = 0; k < 3; k++) {
            const int32_t u = nd.in[k];
            if (u < 0) continue;
            const int32_t lu = (nodes[u].state == NS_LAZY) ? nodes[u].local : leaf_local(u);
            g.info[lu].uses++;
        }
    }
    const int32_t T = red ? targets[0] : -1;
    int32_t weight_local = -1;
    if (red && red->weight >= 0) { weight_local = leaf_local(red->weight); g.info[weight_local].uses++; }
    int32_t target_local = -1;
    if (red) {
        target_local = (nodes[T].state == NS_LAZY) ? nodes[T].local : leaf_local(T);
        g.info[target_local].uses++;   // the reduction epilogue reads it from acc
    }

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
        f.rem = f.uses;
    }
    if (!red) for (int32_t t : targets) if (nodes[t].state == NS_LAZY) g.info[nodes[t].local].store = true;
    for (size_t L = n_cone; L < g.info.size(); L++) g.info[L].rem = g.info[L].uses;

    // ---- 3. emit ----
    g.begin_kernel();
    for (int32_t L = 0; L < n_cone; L++) {
        // margins: one node emits < 16 instructions / < 8 new pointers; a cut spills at most TAPE_REGS + 1 live values
        if ((int)g.ins.size() + 48 > TAPE_MAX_INSTR || (int)g.ptrs.size() + 28 > TAPE_MAX_PTRS) g.cut();
        g.emit_node(L);
    }
    if (red) {
        g.set_pins(target_local, weight_local);
        if (g.acc_owner != target_local) g.take_acc(target_local, 1);
        g.launch(red->mode, red->param, weight_local);
    } else {
        g.launch(RM_NONE, 0.0, -1);
    }

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

void Runtime::reduce(int32_t idx, const ReduceSpec& spec_in, double out[3]) {
    require_init();
    ReduceSpec spec = spec_in;
    if (spec.weight >= 0) materialize(spec.weight);
    if (nodes[idx].n == 0) { out[0] = 0.0; out[1] = NAN; out[2] = NAN; return; }
    std::vector<int32_t> t{idx};
    run_cone(t, &spec);
    FMC_CUDA(cudaMemcpyAsync(h_result, d_result, sizeof(double) * 4, cudaMemcpyDeviceToHost, stream));
    FMC_CUDA(cudaStreamSynchronize(stream));
    stats.d2h += 32;
    out[0] = h_result[0]; out[1] = h_result[1]; out[2] = h_result[2];
}

}  // namespace fmc
