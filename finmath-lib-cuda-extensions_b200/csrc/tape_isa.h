// tape_isa.h — instruction set of the op-tape interpreter (shared by the host code generator and the kernel).
//
// Machine model. One WARP interprets the tape for one chunk of 32 * E consecutive paths at a time, E = 16, 8 or 4 elements
// per lane chosen per launch (TapeHeader::elems; lane l owns elements 128g+4l..128g+4l+3, g = 0..E/4-1, of the chunk, i.e.
// E/4 groups of 128 bits); warps are fully independent of each other (no block barrier on the elementwise path):
//   acc            accumulator, E real registers per lane
//   p              one predicate per element (for choose)
//   slot[0..S-1]   per-warp shared-memory tiles of 128 * E bytes each.
//                  slots [0, n_ring)         "ring": destinations of TMA bulk copies (cp.async.bulk) of leaf-vector
//                                            chunks, each guarded by its own mbarrier; the code generator issues the
//                                            T_LOAD of a chunk as early as the slot is free, so the copy overlaps the
//                                            interpretation of earlier instructions (software prefetch, depth n_ring).
//                                            The prefetch also crosses chunk boundaries: the first instructions of
//                                            the tape (the "prologue", n_prologue T_LOADs) fill the ring before a
//                                            warp's first chunk; the body then re-arms each of those slots with
//                                            T_LOADN (same leaf, the warp's NEXT chunk) as soon as the slot's last
//                                            occupant has been consumed, so every later chunk starts with its first
//                                            leaves already in flight or landed.
//                  slots [n_ring, S)         register file for intermediate values (T_STR writes, *_S reads)
// Instruction word (8 bytes): x = op | slot_byte_offset (slot * 128 * E: the low log2(128 E) bits are the opcode);
//                             y = float immediate bits | pointer-table index.
// Multi-word instructions (T_MULADD_II, ..., T_RATIO) take the y of the following word(s) as further immediates.
// Binary opcodes come in three flavours: _I (operand = immediate), _S (operand = slot), _W (operand = ring slot
// whose TMA copy has not been waited for yet: wait on its mbarrier, then as _S).
// Every primitive rounds once, exactly like the Java float code; the compound RandomVariable ops (accrue, discount,
// addProduct) exist as single instructions but still round after every elementary operation (no FMA).
//
// The opcode numbering is the index into the interpreter's branch-target table (tape_interp.cuh): keep both in sync.
#pragma once
#include <stdint.h>

namespace fmc {

// chunk geometry: E elements per lane (16, 8 or 4), chosen per launch
constexpr int TAPE_E_MAX = 16;
constexpr int TAPE_WARPS = 4;                          // warps per CTA (default)
constexpr int TAPE_MAX_WARPS = 32;                     // ... upper bound (block reduction scratch; codegen.cpp also caps it by the register file: 65536 / (32 * registers per thread))
constexpr bool tape_valid_elems(int e) { return e == 16 || e == 8 || e == 4; }
constexpr int tape_chunk(int elems) { return 32 * elems; }                 // paths per warp iteration: 512 / 256 / 128
constexpr int tape_slot_bytes(int elems) { return 128 * elems; }           // 2 KB / 1 KB / 512 B
constexpr int tape_slot_shift(int elems) { return elems == 16 ? 11 : elems == 8 ? 10 : 9; }
constexpr int TAPE_MAX_RING = 16;                      // ring slots per warp (mbarriers per warp)
constexpr int TAPE_REGS = 16;                          // register-file slots the code generator may use
constexpr int TAPE_MAX_INSTR = 6142;            // a window of several time steps in one launch (codegen.cpp); single-step kernels stay below 2046
constexpr int TAPE_MAX_PTRS = 768;

// binary ops: opcode = T_BIN0 + 3 * k + {0: _I, 1: _S, 2: _W}, k = position in this list
#define FMC_TAPE_BINOPS(X) X(MOV) X(ADD) X(SUB) X(BUS) X(MUL) X(DIV) X(VID) X(MIN) X(MAX) X(SEL) X(ADDPROD) X(ACCRUE) X(DISCOUNT)

enum TapeOp : uint32_t {
    T_END = 0,       // slot operand (optional, y != 0): second operand of a weighted reduction
    T_LOAD = 1,      // ring slot <- ptrs[y][chunk]          (TMA bulk copy, completes on the slot's mbarrier)
    T_WAIT = 2,      // wait for the ring slot's copy
    T_STG = 3,       // ptrs[y][chunk] = acc
    T_STGS = 4,      // ptrs[y][chunk] = slot
    T_STR = 5,       // slot = acc
    T_SETP = 6,      // p = acc >= 0
    T_SQR = 7,       // acc = acc * acc
    T_SQRT = 8, T_EXP = 9, T_LOG = 10, T_SIN = 11, T_COS = 12, T_ABS = 13, T_INV = 14, T_ISNAN = 15,
    T_POW = 16,      // acc = (float) pow((double)acc, (double)imm)
    T_MULADD_II = 17,// acc = acc * imm + imm2              (two words, two roundings)
    T_LOADN = 18,    // ring slot <- ptrs[y][the next chunk that will use this slot set]   (cross-chunk prefetch; no-op near the end)
    T_ACCUM_S = 19,  // acc = acc + slot; slot = acc        (register-file slot)
#define FMC_X(NAME) T_##NAME##_I, T_##NAME##_S, T_##NAME##_W,
    FMC_TAPE_BINOPS(FMC_X)
#undef FMC_X
    T_ADDMUL_II,     // acc = (acc + imm) * imm2            (two words, two roundings; SUB_I a is ADD_I -a exactly)
    T_ADDAFF_S,      // acc = acc + (slot + imm) * imm2     (two words, three roundings: a payoff added to a running value)
    T_ADDAFF_W,      //   ... on a ring slot whose copy has not been waited for yet
    T_MULADDMUL,     // acc = (acc * imm + imm2) * imm3                      (three words)
    T_RATIO,         // acc = (imm3 / (acc * imm + imm2)) * imm4             (four words: p / (1 + L p) * sigma of an LMM drift term)
    T_ADDAFFDISC_S,  // acc = (acc + (slot + imm) * imm2) / (1 + slot * imm3) (three words: one swap period of a swaption)
    T_ADDAFFDISC_W,
    T_ADDAFFDISC_SL, // the same (four words), then the ring slot is re-armed with this chunk of ptrs[y of the fourth word]: the slot's
    T_ADDAFFDISC_WL, //   last use and the T_LOAD that follows it in one dispatch
    T_RATIOACC_S,    // acc = (imm3 / (slot * imm + imm2)) * imm4 + slot2; slot2 = acc   (four words; the fourth word's x names slot2, a
    T_RATIOACC_W,    //   register-file slot: one LMM drift term added to its running sum = MOV ; RATIO ; ACCUM_S)
    T_AXPYST_S,      // acc = (acc * imm + imm2) * imm3 + slot + slot2 * imm4; ptrs[p] = acc; then slot is re-armed with ptrs[q] unless
                     //   q == 0xffffffff   (six words: imm, imm2, imm3, imm4 | slot2, p, q: one LMM state update = MULADDMUL ; ADD_S ;
                     //   ADDPROD_S ; STG and the T_LOAD behind the slot's last use)
    T_RATIOACC_A,    // slot = acc; acc = (imm3 / (acc * imm + imm2)) * imm4 + slot2; slot2 = acc   (four words like T_RATIOACC_S; slot and slot2
                     //   are register-file slots: the drift term of a state that an earlier instruction of the same kernel left in acc — a
                     //   time step inside a window, codegen.cpp — parked for the state update that follows = STR ; RATIO ; ACCUM_S)
    T_NUM_OPS
};
constexpr uint32_t T_BIN0 = 20;
// operand semantics of the binary ops (b = operand, s = immediate):
//   MOV acc = b      ADD acc + b     SUB acc - b     BUS b - acc     MUL acc * b     DIV acc / b     VID b / acc
//   MIN Math.min(acc, b)   MAX Math.max(acc, b)   (NaN propagating, -0 < +0)      SEL p ? acc : b
//   ADDPROD  acc + b * s        ACCRUE  acc * (1 + b * s)        DISCOUNT  acc / (1 + b * s)     (_S/_W only)
static_assert(T_MOV_I == T_BIN0, "binary opcode layout");
static_assert(T_ADDMUL_II == T_BIN0 + 39 && T_ADDAFF_W == T_BIN0 + 41 && T_ADDAFFDISC_W == T_BIN0 + 45 && T_AXPYST_S == T_BIN0 + 50 && T_RATIOACC_A == T_BIN0 + 51 && T_NUM_OPS == T_BIN0 + 52, "binary opcode layout");

enum ReduceMode : int {
    RM_NONE = 0,
    RM_SUM = 1,       // sum(acc)                      -> result[0]
    RM_MOMENTS = 2,   // (count, mean, M2) of acc       -> result[0..2]
    RM_MIN = 3, RM_MAX = 4,
    RM_DOT = 5,       // sum((double)acc * (double)b)
    RM_WSQ = 6        // sum(((double)acc - param)^2 * (double)b)
};

// exchange step of a reduction over a path-sharded vector (reduce_common.cuh): tables[r] = rank r's exchange table as mapped
// in this process (cudaIpc), slot [ticket % XSLOTS][rank] = {count, value, M2, ticket}
constexpr int XMAX_RANKS = 8;
constexpr int XSLOTS = 64;
// host_table != nullptr: the ranks meet in a table in shared HOST memory instead (comm.cpp): the kernel only publishes its own
// partial there and returns, the hosts merge — no kernel ever waits for another rank
struct Exchange { double* tables[XMAX_RANKS]; double* host_table; int rank, nranks; };   // nranks <= 1: no exchange

struct TapeInstr { uint32_t x, y; };

inline TapeInstr enc_idx(uint32_t op, uint32_t slot, uint32_t idx, int shift) { return TapeInstr{ op | (slot << shift), idx }; }

// Everything a launch needs except the two tables.
struct TapeHeader {
    long long n;              // elements per vector
    int elems;                // chunk geometry: elements per lane (16, 8 or 4)
    int n_instr;              // instructions including the final T_END (two more padding words follow)
    int n_prologue;           // leading T_LOADs, closed by a T_END: run once per slot set before the warp's first chunks;
                              // the body starts at instr[n_prologue + 1]
    int n_sets;               // slot sets per warp (cross-chunk prefetch depth), see tape_kernel.cu
    int n_ptrs;
    int n_ring;               // ring slots per warp
    int n_slots;              // ring + register-file slots per warp
    int reduce_mode;
    double reduce_param;
    double* partials;         // [gridDim.x][4]
    unsigned int* counter;    // last-block ticket
    double* result;           // [4]
    double* host_result;      // [4] the same result through mapped pinned host memory ([3] = ticket), or nullptr
    double ticket;            // written to host_result[3] last: the host spins on it instead of a copy + stream sync
    Exchange xchg;            // peer tables of a path-sharded run (reduce_common.cuh)
};

// Host-side description of one launch (code generator -> launch_tape).
struct TapeParams : TapeHeader {
    float* ptrs[TAPE_MAX_PTRS];
    TapeInstr instr[TAPE_MAX_INSTR + 3];   // + closing T_END + two padding words (the interpreter prefetches two ahead)
};

// Kernel arguments. Kernel parameters beyond 4 KB cost ~10 us per launch on the device and ~2 us on the host (measured:
// a trivial fused reduction took 32 us with the 20 KB TapeParams by value, 22 us below 4 KB), so a tape that fits travels
// inline in a 4 KB argument block and a longer one through a ring of device buffers filled by cudaMemcpyAsync.
constexpr int TAPE_INLINE_PTRS = 64;
constexpr int TAPE_INLINE_INSTR = 426;     // words including the closing T_END and the two padding words
struct TapeArgsInline { TapeHeader h; float* ptrs[TAPE_INLINE_PTRS]; TapeInstr instr[TAPE_INLINE_INSTR]; };
struct TapeArgsDev { TapeHeader h; float* const* ptrs; const TapeInstr* instr; };
// ... and the launch latency still grows with the size of the argument block below 4 KB: the tape of a short valuation (a swaption
// of a few periods: a few dozen words) travels in a block of under 1 KB
constexpr int TAPE_SMALL_PTRS = 16;
constexpr int TAPE_SMALL_INSTR = 80;
struct TapeArgsSmall { TapeHeader h; float* ptrs[TAPE_SMALL_PTRS]; TapeInstr instr[TAPE_SMALL_INSTR]; };
static_assert(sizeof(TapeArgsSmall) <= 1024, "the small argument block");
static_assert(sizeof(TapeArgsInline) <= 4096, "the inline argument block must stay within the 4 KB fast path");
constexpr bool tape_fits_inline(int n_ptrs, int n_instr) { return n_instr + 2 <= TAPE_INLINE_INSTR && n_ptrs <= TAPE_INLINE_PTRS; }   // which of the two a launch gets
template <typename ARGS> struct tape_args_in_global { static constexpr bool value = false; };
template <> struct tape_args_in_global<TapeArgsDev> { static constexpr bool value = true; };

}  // namespace fmc
