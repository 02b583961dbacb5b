// comm.cpp — multi-GPU exchange step. One process per GPU, each owning a contiguous path slice of every vector;
// the only data that crosses NVLink are reduction partials and regression normal equations (a few doubles),
// exchanged with ncclAllReduce on the compute stream. The reference has no multi-device code at all
// (one device chosen by a system property, RandomVariableCuda.java:161,177).
//
// NCCL is loaded with dlopen so that libfmcuda.so has no link-time dependency on it (a Java or Python host that
// already loaded libnccl.so.2 shares that copy).
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "runtime.h"

namespace fmc {

namespace {

// minimal subset of nccl.h (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef enum { ncclFloat64 = 8 } ncclDataType_t;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
    static NcclApi a;
    if (a.handle) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (a.handle) break; }
    if (!a.handle) fail(FMC_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.AllGather)
        fail(FMC_ERR_COMM, "libnccl.so.2 lacks a required symbol");
    return a;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) {
        NcclApi& a = api();
        fail(FMC_ERR_COMM, "%s failed: %s", what, a.GetErrorString ? a.GetErrorString(r) : "nccl error");
    }
}

}  // namespace

void comm_get_unique_id(char* id) {
    static_assert(sizeof(ncclUniqueId) == FMC_UNIQUE_ID_BYTES, "unique id size");
    ncclUniqueId uid;
    check(api().GetUniqueId(&uid), "ncclGetUniqueId");
    std::memcpy(id, &uid, sizeof(uid));
}

// Peer tables for the in-kernel exchange: every rank allocates XSLOTS * XMAX_RANKS * 4 doubles, the cudaIpc handles go
// round with one ncclAllGather, every rank maps the others' tables. Falls back to the NCCL exchange if a peer cannot be mapped.
static void setup_peer_tables(Runtime& rt) {
    rt.p2p_ready = false;
    if (!rt.opt.p2p_reduce || rt.comm_size > XMAX_RANKS) return;
    const size_t bytes = sizeof(double) * XSLOTS * XMAX_RANKS * 4;
    FMC_CUDA(cudaMalloc(&rt.xtable, bytes));
    FMC_CUDA(cudaMemsetAsync(rt.xtable, 0, bytes, rt.stream));     // on the compute stream: ordered before the first reduction kernel
    cudaIpcMemHandle_t mine;
    FMC_CUDA(cudaIpcGetMemHandle(&mine, rt.xtable));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    const int R = rt.comm_size;
    char* d_handles = nullptr;
    FMC_CUDA(cudaMalloc(&d_handles, 64 * (size_t)(R + 1)));
    FMC_CUDA(cudaMemcpyAsync(d_handles + 64 * (size_t)R, &mine, 64, cudaMemcpyHostToDevice, rt.stream));
    check(api().AllGather(d_handles + 64 * (size_t)R, d_handles, 8, ncclFloat64, (ncclComm_t)rt.nccl_comm, rt.stream), "ncclAllGather(ipc handles)");
    std::vector<cudaIpcMemHandle_t> all((size_t)R);
    FMC_CUDA(cudaMemcpyAsync(all.data(), d_handles, 64 * (size_t)R, cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    cudaFree(d_handles);
    bool ok = true;
    // the kernels of this exchange wait for one another: every rank needs a GPU of its own (two ranks that time-slice one device
    // would spin on a peer kernel that cannot run). Compare the devices' UUIDs.
    {
        double* d_uuid = nullptr;
        FMC_CUDA(cudaMalloc(&d_uuid, 16 * (size_t)(R + 1)));
        FMC_CUDA(cudaMemcpyAsync((char*)d_uuid + 16 * (size_t)R, &rt.prop.uuid, 16, cudaMemcpyHostToDevice, rt.stream));
        check(api().AllGather((char*)d_uuid + 16 * (size_t)R, d_uuid, 2, ncclFloat64, (ncclComm_t)rt.nccl_comm, rt.stream), "ncclAllGather(device uuids)");
        std::vector<char> uu(16 * (size_t)R);
        FMC_CUDA(cudaMemcpyAsync(uu.data(), d_uuid, 16 * (size_t)R, cudaMemcpyDeviceToHost, rt.stream));
        FMC_CUDA(cudaStreamSynchronize(rt.stream));
        cudaFree(d_uuid);
        for (int a = 0; a < R && ok; a++)
            for (int b = a + 1; b < R; b++)
                if (std::memcmp(&uu[16 * (size_t)a], &uu[16 * (size_t)b], 16) == 0) { ok = false; break; }
    }
    for (int r = 0; r < R && ok; r++) {
        if (r == rt.comm_rank) { rt.peer_tables[r] = rt.xtable; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        rt.peer_tables[r] = (double*)p;
    }
    // everybody must agree: one rank that cannot map a peer sends all of them to the NCCL path
    double flag = ok ? 0.0 : 1.0;
    FMC_CUDA(cudaMemcpyAsync(rt.d_result + 200, &flag, sizeof(double), cudaMemcpyHostToDevice, rt.stream));
    rt.allreduce_sum(rt.d_result + 200, 1);
    FMC_CUDA(cudaMemcpyAsync(&flag, rt.d_result + 200, sizeof(double), cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    rt.p2p_ready = (flag == 0.0);
    rt.xticket = 0.0;
}

// Exchange table in shared HOST memory: one POSIX shm segment per communicator (its name is a hash of the NCCL id every rank
// holds), mapped by every rank and registered with CUDA so that the rank's kernels can store into it. Works whatever GPUs the
// ranks sit on (also two ranks on one device, where a rendez-vous of kernels would dead-lock).
static void setup_host_table(Runtime& rt, const char* id) {
    rt.xhost_ready = false;
    uint64_t hsh = 1469598103934665603ull;
    for (int i = 0; i < FMC_UNIQUE_ID_BYTES; i++) { hsh ^= (unsigned char)id[i]; hsh *= 1099511628211ull; }
    char name[64];
    std::snprintf(name, sizeof(name), "/fmc_xchg_%016llx", (unsigned long long)hsh);
    const size_t bytes = sizeof(double) * XSLOTS * XMAX_RANKS * 4;
    bool ok = rt.comm_size <= XMAX_RANKS;
    int fd = -1;
    void* map = MAP_FAILED;
    if (ok) {
        if (rt.comm_rank == 0) {
            shm_unlink(name);
            fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd >= 0 && ftruncate(fd, (off_t)bytes) != 0) { close(fd); fd = -1; }
        }
    }
    // rank 0 has created the segment before anybody tries to open it
    double flag = (rt.comm_rank == 0 && fd < 0 && ok) ? 1.0 : 0.0;
    FMC_CUDA(cudaMemcpyAsync(rt.d_result + 200, &flag, sizeof(double), cudaMemcpyHostToDevice, rt.stream));
    rt.allreduce_sum(rt.d_result + 200, 1);
    FMC_CUDA(cudaMemcpyAsync(&flag, rt.d_result + 200, sizeof(double), cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    if (flag != 0.0) ok = false;
    if (ok && rt.comm_rank != 0) fd = shm_open(name, O_RDWR, 0600);
    if (ok && fd >= 0) map = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (fd >= 0) close(fd);
    bool mine = ok && map != MAP_FAILED;
    if (mine && rt.comm_rank == 0) std::memset(map, 0, bytes);
    if (mine) {
        if (cudaHostRegister(map, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); mine = false; }
    }
    void* dev = nullptr;
    if (mine && cudaHostGetDevicePointer(&dev, map, 0) != cudaSuccess) { cudaGetLastError(); cudaHostUnregister(map); mine = false; }
    // everybody must agree (and rank 0's zero fill is complete before anybody publishes)
    flag = mine ? 0.0 : 1.0;
    FMC_CUDA(cudaMemcpyAsync(rt.d_result + 200, &flag, sizeof(double), cudaMemcpyHostToDevice, rt.stream));
    rt.allreduce_sum(rt.d_result + 200, 1);
    FMC_CUDA(cudaMemcpyAsync(&flag, rt.d_result + 200, sizeof(double), cudaMemcpyDeviceToHost, rt.stream));
    FMC_CUDA(cudaStreamSynchronize(rt.stream));
    if (rt.comm_rank == 0) shm_unlink(name);             // every rank holds its mapping: the name can go
    if (flag != 0.0) {
        if (mine) cudaHostUnregister(map);
        if (map != MAP_FAILED) munmap(map, bytes);
        return;
    }
    rt.xhost = (double*)map; rt.xhost_dev = (double*)dev; rt.xhost_bytes = bytes;
    rt.xhost_ready = true;
    rt.xticket = 0.0;
}

static void release_host_table(Runtime& rt) {
    if (rt.xhost) { cudaHostUnregister(rt.xhost); munmap(rt.xhost, rt.xhost_bytes); }
    rt.xhost = nullptr; rt.xhost_dev = nullptr; rt.xhost_bytes = 0; rt.xhost_ready = false;
}

static void release_peer_tables(Runtime& rt) {
    for (int r = 0; r < XMAX_RANKS; r++) {
        if (rt.peer_tables[r] && rt.peer_tables[r] != rt.xtable) cudaIpcCloseMemHandle(rt.peer_tables[r]);
        rt.peer_tables[r] = nullptr;
    }
    if (rt.xtable) { cudaFree(rt.xtable); rt.xtable = nullptr; }
    rt.p2p_ready = false;
}

void comm_init(Runtime& rt, int rank, int nranks, const char* id) {
    if (nranks < 1 || rank < 0 || rank >= nranks) fail(FMC_ERR_INVALID, "bad rank %d / %d", rank, nranks);
    comm_destroy(rt);
    if (nranks == 1) { rt.comm_rank = 0; rt.comm_size = 1; return; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    check(api().CommInitRank(&c, nranks, uid, rank), "ncclCommInitRank");
    rt.nccl_comm = c; rt.comm_rank = rank; rt.comm_size = nranks;
    setup_host_table(rt, id);
    setup_peer_tables(rt);
}

void comm_destroy(Runtime& rt) {
    if (rt.nccl_comm) {
        if (rt.stream) cudaStreamSynchronize(rt.stream);
        release_peer_tables(rt);
        release_host_table(rt);
        api().CommDestroy((ncclComm_t)rt.nccl_comm);
        rt.nccl_comm = nullptr;
    }
    rt.comm_rank = 0; rt.comm_size = 1;
}

void Runtime::fill_exchange(Exchange& x, double* ticket) {
    for (int r = 0; r < XMAX_RANKS; r++) x.tables[r] = nullptr;
    x.host_table = nullptr;
    x.rank = 0; x.nranks = 1;
    if (use_xhost()) {
        x.host_table = xhost_dev;
        x.rank = comm_rank; x.nranks = comm_size;
        *ticket = (xticket += 1.0);
    } else if (use_p2p()) {
        for (int r = 0; r < comm_size; r++) x.tables[r] = peer_tables[r];
        x.rank = comm_rank; x.nranks = comm_size;
        *ticket = (xticket += 1.0);
    }
}

void Runtime::allreduce_sum(double* dev, int count) {
    if (comm_size <= 1) return;
    check(api().AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)nccl_comm, stream), "ncclAllReduce");
    stats.n_kernels++;
}

void Runtime::allgather(const double* dev_send, double* dev_recv, int count_per_rank) {
    if (comm_size <= 1) return;
    check(api().AllGather(dev_send, dev_recv, (size_t)count_per_rank, ncclFloat64, (ncclComm_t)nccl_comm, stream), "ncclAllGather");
    stats.n_kernels++;
}

void Runtime::allreduce_minmax(double* dev, int count, bool is_max) {
    if (comm_size <= 1) return;
    check(api().AllReduce(dev, dev, (size_t)count, ncclFloat64, is_max ? ncclMax : ncclMin, (ncclComm_t)nccl_comm, stream), "ncclAllReduce");
    stats.n_kernels++;
}

}  // namespace fmc
