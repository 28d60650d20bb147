// comm.cu — per-context NCCL communicator (one process per GPU), loaded with dlopen so that the
// library has no link-time NCCL dependency and shares whatever libnccl.so.2 the host process
// already mapped (PyTorch bundles its own).  New functionality: the reference is single-device
// (SURVEY.md 2.2); the collective is the all-gather of source records over NVLink each step.
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace pcuda {

// Minimal NCCL surface (nccl.h, NCCL 2.x ABI).
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // 0 == ncclSuccess
enum { ncclInt8 = 0 };

struct Nccl {
    void *handle = nullptr;
    ncclComm_t comm = nullptr;
    int world = 0, rank = 0;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static int nccl_load(pcuda_ctx *ctx) {
    if (ctx->nccl) return PCUDA_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(ctx, PCUDA_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    Nccl *n = new Nccl();
    n->handle = h;
    n->GetUniqueId = (decltype(n->GetUniqueId))dlsym(h, "ncclGetUniqueId");
    n->CommInitRank = (decltype(n->CommInitRank))dlsym(h, "ncclCommInitRank");
    n->CommDestroy = (decltype(n->CommDestroy))dlsym(h, "ncclCommDestroy");
    n->AllGather = (decltype(n->AllGather))dlsym(h, "ncclAllGather");
    n->GetErrorString = (decltype(n->GetErrorString))dlsym(h, "ncclGetErrorString");
    n->Send = (decltype(n->Send))dlsym(h, "ncclSend");  // optional: only the all-to-all needs them
    n->Recv = (decltype(n->Recv))dlsym(h, "ncclRecv");
    n->GroupStart = (decltype(n->GroupStart))dlsym(h, "ncclGroupStart");
    n->GroupEnd = (decltype(n->GroupEnd))dlsym(h, "ncclGroupEnd");
    if (!n->GetUniqueId || !n->CommInitRank || !n->CommDestroy || !n->AllGather) {
        delete n;
        return fail(ctx, PCUDA_ERR_NCCL, "libnccl.so.2 lacks required symbols");
    }
    ctx->nccl = n;
    return PCUDA_OK;
}

void nccl_free(pcuda_ctx *ctx) {
    if (!ctx->nccl) return;
    if (ctx->nccl->comm) ctx->nccl->CommDestroy(ctx->nccl->comm);
    delete ctx->nccl;  // the dlopen handle is intentionally kept: other users may share it
    ctx->nccl = nullptr;
}

void nccl_world(const pcuda_ctx *ctx, int *world, int *rank) {
    const bool on = ctx->nccl && ctx->nccl->comm;
    *world = on ? ctx->nccl->world : 1;
    *rank = on ? ctx->nccl->rank : 0;
}

static int nccl_fail(pcuda_ctx *ctx, const char *what, ncclResult_t r) {
    return fail(ctx, PCUDA_ERR_NCCL, "%s failed: %s (%d)", what,
                ctx->nccl && ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?", r);
}

bool nccl_has_p2p(const pcuda_ctx *ctx) {
    const Nccl *n = ctx->nccl;
    return n && n->comm && n->Send && n->Recv && n->GroupStart && n->GroupEnd;
}

// Variable all-to-all on the context stream: rank o gets send_bytes[o] bytes from d_send +
// send_off[o] and delivers recv_bytes[o] bytes to d_recv + recv_off[o]; one grouped batch of
// ncclSend / ncclRecv.  The own share is a device-to-device copy.  Zero-byte pairs are skipped
// (both sides know the whole matrix, so they skip the same pairs).
int nccl_alltoallv(pcuda_ctx *ctx, const void *d_send, const size_t *send_off, const size_t *send_bytes,
                   void *d_recv, const size_t *recv_off, const size_t *recv_bytes) {
    if (!nccl_has_p2p(ctx)) return fail(ctx, PCUDA_ERR_NCCL, "ncclSend / ncclRecv are not available");
    Nccl *n = ctx->nccl;
    const char *s = static_cast<const char *>(d_send);
    char *r = static_cast<char *>(d_recv);
    const int me = n->rank;
    if (send_bytes[me] != recv_bytes[me]) return fail(ctx, PCUDA_ERR_NCCL, "all-to-all: own share mismatch");
    if (send_bytes[me]) {
        cudaError_t e = cudaMemcpyAsync(r + recv_off[me], s + send_off[me], send_bytes[me],
                                        cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, PCUDA_ERR_CUDA, "all-to-all: own copy: %s", cudaGetErrorString(e));
    }
    ncclResult_t rc = n->GroupStart();
    if (rc) return nccl_fail(ctx, "ncclGroupStart", rc);
    ncclResult_t first = 0;
    for (int o = 0; o < n->world; ++o) {
        if (o == me) continue;
        if (send_bytes[o]) {
            rc = n->Send(s + send_off[o], send_bytes[o], ncclInt8, o, n->comm, ctx->stream);
            if (rc && !first) first = rc;
        }
        if (recv_bytes[o]) {
            rc = n->Recv(r + recv_off[o], recv_bytes[o], ncclInt8, o, n->comm, ctx->stream);
            if (rc && !first) first = rc;
        }
    }
    rc = n->GroupEnd();
    if (first) return nccl_fail(ctx, "ncclSend/ncclRecv", first);
    if (rc) return nccl_fail(ctx, "ncclGroupEnd", rc);
    return PCUDA_OK;
}

}  // namespace pcuda

using namespace pcuda;

extern "C" {

int pcuda_comm_unique_id(pcuda_ctx *ctx, uint8_t id[PCUDA_UNIQUE_ID_BYTES]) {
    if (!ctx || !id) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    PCUDA_TRY(nccl_load(ctx));
    ncclUniqueId uid;
    ncclResult_t r = ctx->nccl->GetUniqueId(&uid);
    if (r) return nccl_fail(ctx, "ncclGetUniqueId", r);
    static_assert(sizeof(uid) == PCUDA_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &uid, sizeof uid);
    return PCUDA_OK;
}

int pcuda_comm_init(pcuda_ctx *ctx, const uint8_t id[PCUDA_UNIQUE_ID_BYTES], int world_size,
                    int rank) {
    if (!ctx || !id || world_size < 1 || rank < 0 || rank >= world_size)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "bad communicator arguments");
    PCUDA_TRY(nccl_load(ctx));
    DeviceGuard guard(ctx->device);
    if (ctx->nccl->comm) {
        ctx->nccl->CommDestroy(ctx->nccl->comm);
        ctx->nccl->comm = nullptr;
    }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclResult_t r = ctx->nccl->CommInitRank(&ctx->nccl->comm, world_size, uid, rank);
    if (r) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl->world = world_size;
    ctx->nccl->rank = rank;
    return PCUDA_OK;
}

int pcuda_comm_destroy(pcuda_ctx *ctx) {
    if (!ctx) return PCUDA_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    nccl_free(ctx);
    return PCUDA_OK;
}

int pcuda_comm_allgather_dev(pcuda_ctx *ctx, const void *d_send, void *d_recv,
                             size_t bytes_per_rank) {
    if (!ctx) return PCUDA_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl || !ctx->nccl->comm)
        return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "pcuda_comm_init has not been called");
    DeviceGuard guard(ctx->device);
    ncclResult_t r = ctx->nccl->AllGather(d_send, d_recv, bytes_per_rank, ncclInt8, ctx->nccl->comm,
                                          ctx->stream);
    if (r) return nccl_fail(ctx, "ncclAllGather", r);
    return PCUDA_OK;
}

}  // extern "C"
