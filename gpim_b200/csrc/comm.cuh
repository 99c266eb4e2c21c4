// comm.cuh -- multi-GPU plumbing of the C ABI: one process per GPU, one NCCL communicator per handle.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, i.e. the copy the process has already loaded -- torch's -- or
// the system one), so libgpgrid.so itself carries no link-time dependency on it and still loads on a box without
// NCCL; only the gpg_comm_* / *_sharded entry points need it.
//
// What travels: rank `root` alone assembles K and factorises it (training / Cholesky are replicas-only); the factor
// cache {theta, X, alpha, scales, info} + {fp16 planes of Linv | Linv} is broadcast over NVLink; every rank predicts
// its tile of X_full rows; one all-gather returns (mean, sd).  On the tcgen05 route the broadcast of the planes is
// PIPELINED by row blocks on the handle's communication stream: the variance GEMM of n-block b only needs rows
// [256 b, 256 b + 256) of Linv, so the first tile of test points starts on the blocks that have landed while the
// rest is still in flight (gpg_predict_sharded).
#pragma once
#include <dlfcn.h>
#include "common.cuh"

namespace comm {

// the slice of nccl.h this file uses (ABI-stable since NCCL 2.0; types restated so that no NCCL header is needed)
struct UniqueId { char internal[128]; };
typedef void *Comm;
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2 };

struct Api {
    void *lib = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

inline Api *api() {
    static Api a;
    static bool tried = false;
    if (tried) return a.lib ? &a : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return nullptr;
    bool ok = true;
    auto sym = [&](const char *n) { void *p = dlsym(a.lib, n); ok = ok && p != nullptr; return p; };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(sym("ncclGetVersion"));
    if (!ok) { dlclose(a.lib); a.lib = nullptr; return nullptr; }
    return &a;
}

#define GPG_NCCL_CHECK(expr)                                                                         \
    do {                                                                                             \
        int _r = (expr);                                                                             \
        if (_r != 0) {                                                                               \
            gpg_set_error("%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #expr, comm::api()->GetErrorString(_r)); \
            return GPG_ECUDA;                                                                        \
        }                                                                                            \
    } while (0)

struct State {
    Comm comm = nullptr;
    int nranks = 1, rank = 0;
    cudaStream_t stream = nullptr;          // non-blocking: every NCCL call of the handle is issued here, in one order
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_small = nullptr;
    std::vector<cudaEvent_t> ev_chunk;      // one per row block of a pipelined broadcast
};

inline int need(gpg_handle_s *h, State **out) {
    State *st = reinterpret_cast<State *>(h->comm);
    if (!st || !st->comm) { gpg_set_error("no communicator on this handle: call gpg_comm_init first"); return GPG_EINVAL; }
    *out = st;
    return GPG_OK;
}

// the communication stream picks up after everything enqueued on `s` so far ...
inline int fork_from(State *st, cudaStream_t s) {
    GPG_CUDA_CHECK(cudaEventRecord(st->ev_in, s));
    GPG_CUDA_CHECK(cudaStreamWaitEvent(st->stream, st->ev_in, 0));
    return GPG_OK;
}
// ... and `s` continues once the communication stream has drained
inline int join_into(State *st, cudaStream_t s) {
    GPG_CUDA_CHECK(cudaEventRecord(st->ev_out, st->stream));
    GPG_CUDA_CHECK(cudaStreamWaitEvent(s, st->ev_out, 0));
    return GPG_OK;
}

inline int bcast_bytes(State *st, void *p, size_t bytes, int root) {
    if (!p || bytes == 0) return GPG_OK;
    GPG_NCCL_CHECK(api()->Broadcast(p, p, bytes, NCCL_UINT8, root, st->comm, st->stream));
    return GPG_OK;
}

}  // namespace comm
