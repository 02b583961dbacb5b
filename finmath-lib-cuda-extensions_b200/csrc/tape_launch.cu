// tape_launch.cu — host side of the op-tape interpreter: picks the kernel compiled for the launch's chunk geometry
// (tape_kernel_e16/8/4.cu, one body: tape_interp.cuh) and hands it the tape either inline in the argument block or through
// a ring of device buffers.
#include <cuda_runtime.h>
#include <cstring>

#include "kernels.h"
#include "tape_isa.h"

namespace fmc {

#define FMC_DECL_GEOMETRY(E)                                                                                                             \
    cudaError_t tape_launch_inline_e##E(int rk, const TapeArgsInline& a, int grid, int threads, size_t smem, cudaStream_t stream);       \
    cudaError_t tape_launch_small_e##E(int rk, const TapeArgsSmall& a, int grid, int threads, size_t smem, cudaStream_t stream);         \
    cudaError_t tape_launch_dev_e##E(int rk, const TapeArgsDev& a, int grid, int threads, size_t smem, cudaStream_t stream);             \
    cudaError_t tape_optin_e##E(int dyn_smem);                                                                                           \
    int tape_occupancy_e##E(int rk, int threads, size_t smem_bytes);
FMC_DECL_GEOMETRY(16) FMC_DECL_GEOMETRY(8) FMC_DECL_GEOMETRY(4)
#undef FMC_DECL_GEOMETRY

size_t tape_smem_bytes(int n_ptrs, int n_instr, int n_slots, int n_sets, int n_warps, int elems) {
    size_t s = (size_t)n_warps * (size_t)n_sets * TAPE_MAX_RING * 8;
    s += ((size_t)n_ptrs * 8 + 15) & ~(size_t)15;
    s = (s + ((size_t)n_instr + 2) * 8 + 127) & ~(size_t)127;
    return s + (size_t)n_warps * (size_t)n_sets * (size_t)n_slots * (size_t)tape_slot_bytes(elems);
}

static int reduce_kind(int mode) {
    switch (mode) {
    case RM_NONE: return 0;
    case RM_SUM: case RM_MIN: case RM_MAX: return 1;
    case RM_MOMENTS: return 2;
    default: return 3;
    }
}

// ring of device buffers for tapes that do not fit the inline argument block (pinned host mirror, one event per slot)
namespace {
constexpr int RING_SLOTS = 32;
constexpr size_t RING_PTR_BYTES = sizeof(float*) * TAPE_MAX_PTRS;
constexpr size_t RING_SLOT_BYTES = RING_PTR_BYTES + sizeof(TapeInstr) * (TAPE_MAX_INSTR + 3);
struct TapeRing {
    char* dev = nullptr; char* host = nullptr;
    cudaEvent_t ev[RING_SLOTS] = {nullptr}; bool used[RING_SLOTS] = {false};
    cudaEvent_t ev_up[RING_SLOTS] = {nullptr};     // the slot's upload (on the copy stream) has finished
    unsigned next = 0;
} g_ring;
}  // namespace

// A long tape goes to the device through the COPY stream as soon as the host has it: the host runs several launches ahead of the
// device, so the copy overlaps the kernels already queued on the compute stream, which only waits for the copy's event (a copy
// on the compute stream itself would sit between two kernels: ~10 us of idle GPU per launch).
cudaError_t launch_tape(const TapeParams& P, int grid, int n_warps, cudaStream_t stream, cudaStream_t copy_stream) {
    const int E = P.elems;
    if (E != 16 && E != 8 && E != 4) return cudaErrorInvalidValue;
    if (n_warps < 1 || n_warps > TAPE_MAX_WARPS) return cudaErrorInvalidValue;
    const size_t smem = tape_smem_bytes(P.n_ptrs, P.n_instr, P.n_slots, P.n_sets, n_warps, E);
    const int rk = reduce_kind(P.reduce_mode);
    const int words = P.n_instr + 2;
    if (words <= TAPE_SMALL_INSTR && P.n_ptrs <= TAPE_SMALL_PTRS) {
        TapeArgsSmall a;
        a.h = static_cast<const TapeHeader&>(P);
        std::memcpy(a.ptrs, P.ptrs, sizeof(float*) * (size_t)P.n_ptrs);
        std::memcpy(a.instr, P.instr, sizeof(TapeInstr) * (size_t)words);
        return E == 16 ? tape_launch_small_e16(rk, a, grid, n_warps * 32, smem, stream)
             : E == 8  ? tape_launch_small_e8(rk, a, grid, n_warps * 32, smem, stream)
                       : tape_launch_small_e4(rk, a, grid, n_warps * 32, smem, stream);
    }
    if (tape_fits_inline(P.n_ptrs, P.n_instr)) {
        TapeArgsInline a;
        a.h = static_cast<const TapeHeader&>(P);
        std::memcpy(a.ptrs, P.ptrs, sizeof(float*) * (size_t)P.n_ptrs);
        std::memcpy(a.instr, P.instr, sizeof(TapeInstr) * (size_t)words);
        return E == 16 ? tape_launch_inline_e16(rk, a, grid, n_warps * 32, smem, stream)
             : E == 8  ? tape_launch_inline_e8(rk, a, grid, n_warps * 32, smem, stream)
                       : tape_launch_inline_e4(rk, a, grid, n_warps * 32, smem, stream);
    }
    if (!g_ring.dev) return cudaErrorNotReady;
    const unsigned slot = g_ring.next++ % RING_SLOTS;
    if (g_ring.used[slot]) { const cudaError_t e = cudaEventSynchronize(g_ring.ev[slot]); if (e != cudaSuccess) return e; }   // host mirror still being read
    char* h = g_ring.host + (size_t)slot * RING_SLOT_BYTES;
    char* d = g_ring.dev + (size_t)slot * RING_SLOT_BYTES;
    std::memcpy(h, P.ptrs, sizeof(float*) * (size_t)P.n_ptrs);
    // pointer table and tape are adjacent in the slot when the table is packed to its real size
    const size_t ptr_bytes = (sizeof(float*) * (size_t)P.n_ptrs + 15) & ~(size_t)15;
    std::memcpy(h + ptr_bytes, P.instr, sizeof(TapeInstr) * (size_t)words);
    cudaStream_t up = copy_stream ? copy_stream : stream;
    cudaError_t e = cudaMemcpyAsync(d, h, ptr_bytes + sizeof(TapeInstr) * (size_t)words, cudaMemcpyHostToDevice, up);
    if (e != cudaSuccess) return e;
    if (up != stream) {
        e = cudaEventRecord(g_ring.ev_up[slot], up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, g_ring.ev_up[slot], 0);
        if (e != cudaSuccess) return e;
    }
    TapeArgsDev a;
    a.h = static_cast<const TapeHeader&>(P);
    a.ptrs = reinterpret_cast<float* const*>(d);
    a.instr = reinterpret_cast<const TapeInstr*>(d + ptr_bytes);
    e = E == 16 ? tape_launch_dev_e16(rk, a, grid, n_warps * 32, smem, stream)
      : E == 8  ? tape_launch_dev_e8(rk, a, grid, n_warps * 32, smem, stream)
                : tape_launch_dev_e4(rk, a, grid, n_warps * 32, smem, stream);
    if (e != cudaSuccess) return e;
    e = cudaEventRecord(g_ring.ev[slot], stream);            // the slot is free again when this kernel has run
    if (e != cudaSuccess) return e;
    g_ring.used[slot] = true;
    return cudaSuccess;
}

cudaError_t tape_kernel_setup(size_t* max_smem_per_cta) {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    const int dyn = optin - 1024;                       // static smem of the reduction
    e = tape_optin_e16(dyn);
    if (e == cudaSuccess) e = tape_optin_e8(dyn);
    if (e == cudaSuccess) e = tape_optin_e4(dyn);
    if (e != cudaSuccess) return e;
    if (max_smem_per_cta) *max_smem_per_cta = (size_t)dyn;
    if (!g_ring.dev) {
        e = cudaMalloc(&g_ring.dev, RING_SLOT_BYTES * RING_SLOTS);
        if (e == cudaSuccess) e = cudaMallocHost(&g_ring.host, RING_SLOT_BYTES * RING_SLOTS);
        for (int i = 0; i < RING_SLOTS && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming); g_ring.used[i] = false;
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ring.ev_up[i], cudaEventDisableTiming);
        }
        g_ring.next = 0;
    }
    return e;
}

void tape_kernel_teardown() {
    if (g_ring.dev) cudaFree(g_ring.dev);
    if (g_ring.host) cudaFreeHost(g_ring.host);
    for (int i = 0; i < RING_SLOTS; i++) {
        if (g_ring.ev[i]) { cudaEventDestroy(g_ring.ev[i]); g_ring.ev[i] = nullptr; }
        if (g_ring.ev_up[i]) { cudaEventDestroy(g_ring.ev_up[i]); g_ring.ev_up[i] = nullptr; }
    }
    g_ring.dev = nullptr; g_ring.host = nullptr;
}

int tape_max_blocks_per_sm(size_t smem_bytes, int reduce_mode, int n_warps, int elems) {
    const int rk = reduce_kind(reduce_mode);
    return elems == 16 ? tape_occupancy_e16(rk, n_warps * 32, smem_bytes)
         : elems == 8  ? tape_occupancy_e8(rk, n_warps * 32, smem_bytes)
                       : tape_occupancy_e4(rk, n_warps * 32, smem_bytes);
}

}  // namespace fmc
