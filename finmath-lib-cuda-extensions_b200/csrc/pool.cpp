// pool.cpp — pooled device allocator + pinned staging. Replaces RandomVariableCuda.DeviceMemoryPool's
// per-size ReferenceQueue recycling, its synchronous cudaMemGetInfo per miss and the System.gc() back-off ladder
// (/root/reference/src/main/java/net/finmath/cuda/montecarlo/RandomVariableCuda.java:280-449).
#include <algorithm>
#include <cstdarg>
#include <cstdio>

#include "runtime.h"

namespace fmc {

// ---- errors ----
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
const char* get_error() { return g_err; }
void fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    throw Fail{code};
}

// ---- pool ----
static constexpr size_t kAlign = 512;                 // keeps every vector 128-bit (and sector) aligned
static constexpr size_t kSlabBytes = 256ull << 20;    // default slab; bigger requests get a dedicated slab

size_t DevicePool::round_size(size_t b) {
    if (b == 0) b = 1;
    return (b + kAlign - 1) / kAlign * kAlign;
}

void* DevicePool::carve(size_t rounded) {
    // first fit in the bump region of an existing slab
    for (size_t s = 0; s < slabs_.size(); s++) {
        Slab& sl = slabs_[s];
        if (!sl.dedicated && sl.base && sl.size - sl.used >= rounded) {
            void* p = sl.base + sl.used;
            sl.used += rounded;
            sl.live++;
            blocks_[p] = Block{rounded, (int)s};
            return p;
        }
    }
    const bool dedicated = rounded > kSlabBytes / 4;
    const size_t slab_size = dedicated ? rounded : kSlabBytes;
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, slab_size);
    if (e != cudaSuccess) {
        cudaGetLastError();
        // under pressure: give cached blocks back and retry once (replaces the gc()/wait ladder RVC:308-342)
        trim();
        e = cudaMalloc(&base, slab_size);
        if (e != cudaSuccess) {
            cudaGetLastError();
            fail(FMC_ERR_OOM, "device out of memory: requested %zu bytes (in use %llu, cached %llu, reserved %llu)", slab_size,
                 (unsigned long long)bytes_in_use, (unsigned long long)bytes_cached, (unsigned long long)bytes_reserved);
        }
    }
    bytes_reserved += slab_size;
    int s = -1;
    for (size_t i = 0; i < slabs_.size(); i++) if (!slabs_[i].base) { s = (int)i; break; }
    if (s < 0) { slabs_.push_back(Slab{}); s = (int)slabs_.size() - 1; }
    slabs_[s] = Slab{(char*)base, slab_size, rounded, 1, dedicated};
    blocks_[base] = Block{rounded, s};
    return base;
}

void* DevicePool::alloc(size_t bytes) {
    const size_t rounded = round_size(bytes);
    n_alloc++;
    void* p = nullptr;
    std::vector<Cached>& v = list_for(rounded);
    if (!v.empty()) {
        const Cached c = v.back();
        v.pop_back();
        p = c.p;
        bytes_cached -= rounded;
        slabs_[c.slab].live++;
        n_reused++;
    } else {
        p = carve(rounded);
    }
    bytes_in_use += rounded;
    high_water = std::max(high_water, bytes_in_use);
    return p;
}

void* DevicePool::alloc_settled(size_t bytes, uint64_t settled_stamp, bool* settled) {
    const size_t rounded = round_size(bytes);
    std::vector<Cached>& v = list_for(rounded);     // oldest frees sit at the front (free() appends, alloc() pops the back)
    for (size_t i = 0; i < v.size() && i < 64; i++) {
        if (v[i].stamp > settled_stamp) continue;
        const Cached c = v[i];
        v.erase(v.begin() + (long)i);
        n_alloc++; n_reused++;
        bytes_cached -= rounded;
        slabs_[c.slab].live++;
        bytes_in_use += rounded;
        high_water = std::max(high_water, bytes_in_use);
        *settled = true;
        return c.p;
    }
    if (v.empty()) {                                 // nothing cached of this size: fresh memory, or the usual OOM ladder
        *settled = true;
        n_alloc++;
        void* p = carve(rounded);
        bytes_in_use += rounded;
        high_water = std::max(high_water, bytes_in_use);
        return p;
    }
    *settled = false;                                // cached blocks exist but all were freed too recently
    return alloc(bytes);
}

void DevicePool::free(void* p) {
    if (!p) return;
    auto it = blocks_.find(p);
    if (it == blocks_.end()) return;
    const size_t rounded = it->second.size;
    slabs_[it->second.slab].live--;
    bytes_in_use -= rounded;
    bytes_cached += rounded;
    list_for(rounded).push_back(Cached{p, it->second.slab, ++free_stamp});
}

void DevicePool::trim() {
    // release every slab that holds no live block; its cached blocks disappear with it
    for (size_t s = 0; s < slabs_.size(); s++) {
        Slab& sl = slabs_[s];
        if (!sl.base || sl.live != 0) continue;
        for (auto& kv : free_) {
            auto& v = kv.second;
            for (size_t i = 0; i < v.size();) {
                if (v[i].slab == (int)s) {
                    bytes_cached -= kv.first;
                    blocks_.erase(v[i].p);
                    v.erase(v.begin() + (long)i);            // keeps the oldest-first order
                } else i++;
            }
        }
        cudaFree(sl.base);
        bytes_reserved -= sl.size;
        sl = Slab{nullptr, 0, 0, 0, false};
    }
}

void DevicePool::purge() { trim(); }

// ---- staging ----
void Staging::ensure(size_t need) {
    if (need <= bytes) return;
    release();
    size_t want = std::max<size_t>(need, 8ull << 20);
    FMC_CUDA(cudaMallocHost(&host, want));
    bytes = want;
}
void Staging::release() {
    if (host) cudaFreeHost(host);
    host = nullptr; bytes = 0;
}

}  // namespace fmc
