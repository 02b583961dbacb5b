// comm.cpp — multi-GPU exchange step. One process per GPU, each owning a contiguous path slice of every vector;
// the only data that crosses NVLink are reduction partials and regression normal equations (a few doubles),
// exchanged with ncclAllReduce on the compute stream. The reference has no multi-device code at all
// (one device chosen by a system property, RandomVariableCuda.java:161,177).
//
// NCCL is loaded with dlopen so that libfmcuda.so has no link-time dependency on it (a Java or Python host that
// already loaded libnccl.so.2 shares that copy).
#include <dlfcn.h>

#include <cstring>

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

void comm_init(Runtime& rt, int rank, int nranks, const char* id) {
    if (nranks < 1 || rank < 0 || rank >= nranks) fail(FMC_ERR_INVALID, "bad rank %d / %d", rank, nranks);
    comm_destroy(rt);
    if (nranks == 1) { rt.comm_rank = 0; rt.comm_size = 1; return; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    check(api().CommInitRank(&c, nranks, uid, rank), "ncclCommInitRank");
    rt.nccl_comm = c; rt.comm_rank = rank; rt.comm_size = nranks;
}

void comm_destroy(Runtime& rt) {
    if (rt.nccl_comm) {
        if (rt.stream) cudaStreamSynchronize(rt.stream);
        api().CommDestroy((ncclComm_t)rt.nccl_comm);
        rt.nccl_comm = nullptr;
    }
    rt.comm_rank = 0; rt.comm_size = 1;
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
