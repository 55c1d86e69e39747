// nccl_dl.h — NCCL bound at run time with dlopen, so libqsb loads (and runs single-GPU) on a
// machine without NCCL and never clashes with the copy torch already loaded.  Only the handful
// of entry points the sharded path needs; declarations follow the public nccl.h ABI (2.x).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace qsb {

typedef struct ncclComm* nccl_comm_t;
typedef struct {
    char internal[128];
} nccl_uid_t;
enum { kNcclSuccess = 0, kNcclInProgress = 7 };
enum { kNcclSum = 0, kNcclProd = 1, kNcclMax = 2, kNcclMin = 3 };
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclUint32 = 3, kNcclInt64 = 4, kNcclUint64 = 5,
       kNcclFloat16 = 6, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_uid_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*CommGetAsyncError)(nccl_comm_t, int*) = nullptr;  // optional
    char why[256] = {0};
    bool ok = false;
};

// Always returns the (static) table; check ->ok, and ->why when it is false.
inline NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {getenv("QSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; i < 3 && !api.handle; ++i)
        if (names[i] && names[i][0]) api.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) {
        const char* e = dlerror();
        snprintf(api.why, sizeof(api.why), "cannot load NCCL (set QSB_NCCL_LIB): %s", e ? e : "?");
        return &api;
    }
#define QSB_NCCL_SYM(field, name)                                                     \
    *(void**)(&api.field) = dlsym(api.handle, name);                                  \
    if (!api.field) {                                                                 \
        snprintf(api.why, sizeof(api.why), "NCCL symbol %s missing", name);           \
        return &api;                                                                  \
    }
    QSB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    QSB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    QSB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    QSB_NCCL_SYM(AllReduce, "ncclAllReduce")
    QSB_NCCL_SYM(AllGather, "ncclAllGather")
    QSB_NCCL_SYM(Send, "ncclSend")
    QSB_NCCL_SYM(Recv, "ncclRecv")
    QSB_NCCL_SYM(GroupStart, "ncclGroupStart")
    QSB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    QSB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef QSB_NCCL_SYM
    *(void**)(&api.CommGetAsyncError) = dlsym(api.handle, "ncclCommGetAsyncError");
    api.ok = true;
    return &api;
}

}  // namespace qsb
