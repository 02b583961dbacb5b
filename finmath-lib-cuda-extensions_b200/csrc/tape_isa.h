// tape_isa.h — instruction set of the op-tape interpreter (shared by the host code generator and the kernel).
//
// Machine model (per path element, E elements per thread):
//   acc          : accumulator, a real register
//   r[0..R-1]    : register file, real registers (statically indexed inside every handler)
//   p            : one predicate (for choose)
// Instruction word (8 bytes): x = op | src_kind << 8 | idx << 16 ; y = float immediate bits.
// An instruction first fetches its operand b from `src` (register j, immediate, leaf vector idx, or acc) and
// then applies op(acc, b). Compound RandomVariable ops (accrue, discount, addProduct, addRatio, subRatio) are
// expanded at record time into these primitives; every primitive rounds once, exactly like the Java float code.
#pragma once
#include <stdint.h>

namespace fmc {

constexpr int TAPE_REGS = 12;         // R: registers the code generator may use (shared-memory register file)
constexpr int TAPE_REGS_FAST = 4;     // tapes that need <= this many run with the register file in real registers
constexpr int TAPE_ELEMS = 8;         // E (two float4 per thread)
constexpr int TAPE_THREADS = 256;
constexpr int TAPE_TILE = TAPE_THREADS * TAPE_ELEMS;   // 2048 elements per block iteration
constexpr int TAPE_MAX_INSTR = 1024;
constexpr int TAPE_MAX_PTRS = 224;

enum TapeOp : uint32_t {
    T_END = 0,      // src (optional): second operand of a weighted reduction
    // ops with operand b
    T_MOV = 1,      // acc = b
    T_ADD = 2,      // acc = acc + b
    T_SUB = 3,      // acc = acc - b
    T_BUS = 4,      // acc = b - acc
    T_MUL = 5,      // acc = acc * b
    T_DIV = 6,      // acc = acc / b
    T_VID = 7,      // acc = b / acc
    T_MIN = 8,      // acc = Math.min(acc, b)   (NaN propagating, -0 < +0)
    T_MAX = 9,      // acc = Math.max(acc, b)
    T_SEL = 10,     // acc = p ? acc : b
    T_STG = 11,     // out[idx_hi] = b        (idx field = output pointer slot; src in y word, see enc_stg)
    T_LAST_WITH_SRC = 11,
    // ops without operand
    T_STR = 12,     // r[idx] = acc
    T_SETP = 13,    // p = acc >= 0
    T_SQRT = 14, T_EXP = 15, T_LOG = 16, T_SIN = 17, T_COS = 18, T_ABS = 19, T_INV = 20, T_ISNAN = 21,
    T_POW = 22,     // acc = (float) pow((double)acc, (double)imm)
    T_NUM_OPS
};

enum TapeSrc : uint32_t {
    S_IMM = 0,      // b = imm
    S_LEAF = 1,     // b = ptrs[idx][i]
    S_ACC = 2,      // b = acc
    S_REG0 = 3      // b = r[src - S_REG0]
};

enum ReduceMode : int {
    RM_NONE = 0,
    RM_SUM = 1,       // sum(acc)                      -> result[0]
    RM_MOMENTS = 2,   // (count, mean, M2) of acc       -> result[0..2]
    RM_MIN = 3, RM_MAX = 4,
    RM_DOT = 5,       // sum((double)acc * (double)b)
    RM_WSQ = 6        // sum(((double)acc - param)^2 * (double)b)
};

struct TapeInstr { uint32_t x, y; };

inline TapeInstr enc(uint32_t op, uint32_t src, uint32_t idx, float imm) {
    union { float f; uint32_t u; } c; c.f = imm;
    return TapeInstr{ op | (src << 8) | (idx << 16), c.u };
}
// STG needs two indices (output slot and source): slot goes into idx, source kind into src (S_ACC or S_REGj).
inline TapeInstr enc_stg(uint32_t out_slot, uint32_t src) { return TapeInstr{ T_STG | (src << 8) | (out_slot << 16), 0u }; }

struct TapeParams {
    long long n;              // elements per vector
    int n_instr;
    int reduce_mode;
    double reduce_param;
    double* partials;         // [gridDim.x][4]
    unsigned int* counter;    // last-block ticket
    double* result;           // [4]
    float* ptrs[TAPE_MAX_PTRS];
    TapeInstr instr[TAPE_MAX_INSTR + 2];
};

}  // namespace fmc
