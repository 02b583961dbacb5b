// reduce_common.cuh — what the two reduction paths (the interpreter's fused epilogue, tape_kernel.cu, and the streaming
// reduction, reduce_kernel.cu) share: the {count, value, M2} partial, its deterministic merge, and the FINISH step of a
// reduction — publication of the result to the host and, for a path-sharded vector, the exchange between the GPUs.
//
// Exchange (one process per GPU). The reference has no multi-device code; a collective library call after
// the kernel would cost a second launch, a device->host copy and a stream synchronisation per getAverage(). Instead the
// LAST BLOCK of the reduction kernel itself does the exchange over NVLink peer memory: every rank owns a small table in
// device memory that the other ranks have mapped (cudaIpc); the block stores its rank's partial and then a ticket into
// the slot [ticket % XSLOTS][rank] of every peer's table (plain stores to peer addresses, __threadfence_system between
// data and ticket), waits until the tickets of all ranks have arrived in its OWN table, merges the partials in rank order
// (the same deterministic merge on every rank) and publishes the global result through mapped pinned host memory, where
// the host spins on the ticket. One launch, no collective call, no copy, no stream synchronisation.
// A rank cannot run ahead of the others by more than one reduction (it needs their partials), so slots are never reused
// too early; a peer that never arrives (crashed process) trips a time-out that publishes an error ticket instead of hanging.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "tape_isa.h"

namespace fmc {

constexpr long long XTIMEOUT_CYCLES = 120000000000ll;    // ~60 s (the in-kernel exchange, option exchange=1, is no longer the default)

struct Part { double c, v, m; };       // count, value (sum | mean | min | max), M2

__device__ __forceinline__ double jmin(double a, double b) {      // java.lang.Math.min: NaN propagating, -0 < +0
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (signbit(a) || signbit(b)) ? -0.0 : 0.0;
    return a < b ? a : b;
}
__device__ __forceinline__ double jmax(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0 && b == 0.0) return (signbit(a) && signbit(b)) ? -0.0 : 0.0;
    return a > b ? a : b;
}
__device__ __forceinline__ Part merge(int mode, Part a, Part b) {
    if (b.c == 0.0) return a;
    if (a.c == 0.0) return b;
    Part r;
    r.c = a.c + b.c;
    r.m = 0.0;
    if (mode == RM_MOMENTS) {          // Chan et al. pairwise update
        const double delta = b.v - a.v;
        const double w = b.c / r.c;
        r.v = a.v + delta * w;
        r.m = a.m + b.m + delta * delta * a.c * w;
    } else if (mode == RM_MIN) r.v = jmin(a.v, b.v);
    else if (mode == RM_MAX) r.v = jmax(a.v, b.v);
    else r.v = a.v + b.v;
    return r;
}
__device__ __forceinline__ Part shfl_down(Part p, int d) {
    Part r;
    r.c = __shfl_down_sync(0xffffffffu, p.c, d);
    r.v = __shfl_down_sync(0xffffffffu, p.v, d);
    r.m = __shfl_down_sync(0xffffffffu, p.m, d);
    return r;
}

// Called by ONE thread of the last block with the rank-local result q. Writes the (global) result to result[0..2] and,
// if host != nullptr, to the mapped host mirror followed by the ticket (negative ticket = exchange timed out).
static __device__ __noinline__ void finish_reduction(int mode, Part q, const Exchange& X, double ticket, double* __restrict__ result, double* host)
{
    bool ok = true;
    if (X.nranks > 1 && X.host_table) {
        // exchange through shared host memory: publish this rank's partial and leave; every rank's HOST waits for the R tickets
        // of the slot and merges them in rank order (Runtime::reduce), so no kernel holds its stream for another rank
        const int slot = (int)((long long)ticket % XSLOTS);
        volatile double* t = X.host_table + ((long long)slot * XMAX_RANKS + X.rank) * 4;
        t[0] = q.c; t[1] = q.v; t[2] = q.m;
        __threadfence_system();
        t[3] = ticket;
        result[0] = q.c; result[1] = q.v; result[2] = q.m;
        return;
    }
    if (X.nranks > 1) {
        const int slot = (int)((long long)ticket % XSLOTS);
        for (int r = 0; r < X.nranks; r++) {                          // my partial into everybody's table (mine included)
            volatile double* t = X.tables[r] + ((long long)slot * XMAX_RANKS + X.rank) * 4;
            t[0] = q.c; t[1] = q.v; t[2] = q.m;
        }
        __threadfence_system();
        for (int r = 0; r < X.nranks; r++) {
            volatile double* t = X.tables[r] + ((long long)slot * XMAX_RANKS + X.rank) * 4;
            t[3] = ticket;
        }
        __threadfence_system();
        const volatile double* mine = X.tables[X.rank] + (long long)slot * XMAX_RANKS * 4;
        const long long t0 = clock64();
        Part g = {0.0, 0.0, 0.0};
        for (int r = 0; r < X.nranks && ok; r++) {                    // rank order: the same merge on every rank
            while (mine[r * 4 + 3] != ticket) {
                if (clock64() - t0 > XTIMEOUT_CYCLES) { ok = false; break; }
            }
            if (!ok) break;
            __threadfence_system();
            Part p = { mine[r * 4 + 0], mine[r * 4 + 1], mine[r * 4 + 2] };
            g = merge(mode, g, p);
        }
        q = g;
    }
    result[0] = q.c; result[1] = q.v; result[2] = q.m;
    if (host && X.nranks <= 1 && mode == RM_SUM) {
        // sums of an unsharded vector (every getAverage of a valuation): the host knows the count; value and ticket travel in ONE
        // 16-byte store to {host[2], host[3]} — they become visible together, no system-wide fence between them (2.4 us per reduction)
        asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" :: "l"(host + 2), "d"(q.v), "d"(ticket) : "memory");
        return;
    }
    if (host) {
        volatile double* h = host;
        h[0] = q.c; h[1] = q.v; h[2] = q.m;
        __threadfence_system();
        h[3] = ok ? ticket : -ticket;
    }
}

}  // namespace fmc
