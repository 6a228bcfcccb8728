// pyh_comm.cuh -- multi-rank transport owned by the C layer: NCCL over NVLink, bound at run time.
//
// Replaces the reference's mpi4py layer: one Isend / Irecv per ghost strip (blocks/ghost.py:169-241), the
// Waitall of Blocks.apply_boundary_condition (blocks/base.py:454-465) and the gather + bcast of the time step
// (solvers/base.py:128-131).  libnccl.so.2 is dlopen'ed (the copy torch already mapped into the process if there
// is one, else the system library), so the shared library has no link-time dependency on NCCL and single-rank
// use never touches it.  Only the handful of entry points below are used; their prototypes are restated here.
#pragma once
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

namespace pyh {

// ---- the slice of the NCCL ABI this layer binds (nccl.h 2.x: stable since 2.7) --------------------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
typedef int ncclResult_t;                       // 0 == ncclSuccess
enum { kNcclMin = 3 };                          // ncclRedOp_t: ncclSum 0, ncclProd 1, ncclMax 2, ncclMin 3
enum { kNcclUint64 = 5, kNcclFloat64 = 8 };     // ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    const char* error = nullptr;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char* names[3] = {getenv("PYH_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.error = "libnccl.so.2 not found (set PYH_NCCL_LIB)"; return api; }
#define PYH_NCCL_SYM(field, sym)                                                  \
    *(void**)(&api.field) = dlsym(api.handle, sym);                               \
    if (!api.field) { api.error = "libnccl: symbol " sym " missing"; return api; }
    PYH_NCCL_SYM(GetVersion, "ncclGetVersion")
    PYH_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PYH_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PYH_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PYH_NCCL_SYM(GroupStart, "ncclGroupStart")
    PYH_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PYH_NCCL_SYM(Send, "ncclSend")
    PYH_NCCL_SYM(Recv, "ncclRecv")
    PYH_NCCL_SYM(AllReduce, "ncclAllReduce")
    PYH_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PYH_NCCL_SYM
    return api;
}

// one point-to-point message of the per-stage strip exchange
struct HaloMsg {
    int peer;
    int key_gid, key_side;    // (source block, source side): names the message on both ends
    long long offset, len;    // doubles, inside the send / recv buffer (slot layout of pyh_halo_slot)
};

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::vector<HaloMsg> sends, recvs;   // each sorted by (peer, key): the k-th send a -> b meets the k-th receive b posts for a
    double* d_send = nullptr;
    double* d_recv = nullptr;
    long long doubles = 0;
};

inline bool msg_less(const HaloMsg& a, const HaloMsg& b) {
    if (a.peer != b.peer) return a.peer < b.peer;
    if (a.key_gid != b.key_gid) return a.key_gid < b.key_gid;
    return a.key_side < b.key_side;
}

}  // namespace pyh
