// comm.cu — per-context NCCL communicator (one process per GPU), loaded with dlopen so that the
// library has no link-time NCCL dependency and shares whatever libnccl.so.2 the host process
// already mapped (PyTorch bundles its own).  New functionality: the reference is single-device
// (SURVEY.md 2.2); the collective is the all-gather of source records over NVLink each step.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace pcuda {

// Minimal NCCL surface (nccl.h, NCCL 2.x ABI).
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;  // 0 == ncclSuccess
enum { ncclInt8 = 0 };

// In-process communicator: `world` contexts of ONE process (one host thread each, same device or
// devices with peer access) exchange data with device-to-device copies ordered by events, behind
// the same all-gather / all-to-all calls as NCCL.  It exists so that every multi-GPU path — sharding,
// partitioned and locally-essential tree builds, result routing — can be run and checked on a box
// with a single GPU (tests/test_local_ranks_gpu.py); it is not a data path for production.
constexpr int LOCAL_MAX = 16;
struct LocalGroup {
    int world = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0, attached = 0;
    unsigned long long generation = 0;
    // what every rank published for the collective in flight
    const char *send[LOCAL_MAX] = {};
    const size_t *send_off[LOCAL_MAX] = {}, *send_bytes[LOCAL_MAX] = {};
    cudaEvent_t ready[LOCAL_MAX] = {}, done[LOCAL_MAX] = {};
    bool broken = false;  // a rank did not show up in time: every later rendezvous fails at once
    bool barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const unsigned long long gen = generation;
        if (++arrived == world) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        if (!cv.wait_for(lk, std::chrono::seconds(30), [&] { return generation != gen || broken; }) || broken) {
            broken = true;  // (a rank that failed before the collective never arrives)
            cv.notify_all();
            return false;
        }
        return true;
    }
};

struct Nccl {
    void *handle = nullptr;
    ncclComm_t comm = nullptr;
    LocalGroup *local = nullptr;  // in-process communicator instead of NCCL
    int world = 0, rank = 0;
    bool poisoned = false;  // a collective call failed half-way on this rank: the communicator is unusable
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
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
    n->CommAbort = (decltype(n->CommAbort))dlsym(h, "ncclCommAbort");
    if (!n->GetUniqueId || !n->CommInitRank || !n->CommDestroy || !n->AllGather) {
        delete n;
        return fail(ctx, PCUDA_ERR_NCCL, "libnccl.so.2 lacks required symbols");
    }
    ctx->nccl = n;
    return PCUDA_OK;
}

void nccl_free(pcuda_ctx *ctx) {
    if (!ctx->nccl) return;
    if (LocalGroup *g = ctx->nccl->local) {
        bool last;
        {
            std::lock_guard<std::mutex> lk(g->m);
            last = --g->attached == 0;
        }
        const int r = ctx->nccl->rank;
        if (g->ready[r]) cudaEventDestroy(g->ready[r]);
        if (g->done[r]) cudaEventDestroy(g->done[r]);
        g->ready[r] = g->done[r] = nullptr;
        if (last) delete g;
        ctx->nccl->local = nullptr;
    }
    if (ctx->nccl->comm) ctx->nccl->CommDestroy(ctx->nccl->comm);
    delete ctx->nccl;  // the dlopen handle is intentionally kept: other users may share it
    ctx->nccl = nullptr;
}

void nccl_world(const pcuda_ctx *ctx, int *world, int *rank) {
    const bool on = ctx->nccl && (ctx->nccl->comm || ctx->nccl->local || ctx->nccl->poisoned);
    *world = on ? ctx->nccl->world : 1;
    *rank = on ? ctx->nccl->rank : 0;
}

// A multi-rank step that fails on one rank between two collectives leaves the other ranks waiting in the
// next one; this rank cannot repair that, but it must not add to it: the communicator is aborted (its
// pending operations are cancelled) and every later call on it fails at once instead of hanging.
void nccl_poison(pcuda_ctx *ctx) {
    Nccl *n = ctx->nccl;
    if (!n || n->poisoned) return;
    n->poisoned = true;
    if (n->local) {
        std::lock_guard<std::mutex> lk(n->local->m);
        n->local->broken = true;
        n->local->cv.notify_all();
    } else if (n->comm && n->CommAbort) {
        n->CommAbort(n->comm);
        n->comm = nullptr;
    }
}

bool nccl_poisoned(const pcuda_ctx *ctx) { return ctx->nccl && ctx->nccl->poisoned; }

static int nccl_fail(pcuda_ctx *ctx, const char *what, ncclResult_t r) {
    return fail(ctx, PCUDA_ERR_NCCL, "%s failed: %s (%d)", what,
                ctx->nccl && ctx->nccl->GetErrorString ? ctx->nccl->GetErrorString(r) : "?", r);
}

bool nccl_has_p2p(const pcuda_ctx *ctx) {
    const Nccl *n = ctx->nccl;
    return n && (n->local || (n->comm && n->Send && n->Recv && n->GroupStart && n->GroupEnd));
}

// In-process all-to-all: every rank publishes its send buffer and an event recorded behind the work
// that fills it, meets the others at a host barrier, copies its shares out of the peers' buffers on
// its own stream (behind their events), and meets them again so that nobody's send buffer is reused
// before every peer has queued its copy (a second event per rank orders that on the device).
static int local_alltoallv(pcuda_ctx *ctx, const void *d_send, const size_t *send_off, const size_t *send_bytes,
                           void *d_recv, const size_t *recv_off, const size_t *recv_bytes) {
    Nccl *n = ctx->nccl;
    LocalGroup *g = n->local;
    const int me = n->rank;
    int status = PCUDA_OK;
    auto check = [&](cudaError_t e) {  // no early return: every rank must reach both barriers
        if (e != cudaSuccess && status == PCUDA_OK)
            status = fail(ctx, PCUDA_ERR_CUDA, "local communicator: %s", cudaGetErrorString(e));
    };
    check(cudaEventRecord(g->ready[me], ctx->stream));
    g->send[me] = static_cast<const char *>(d_send);
    g->send_off[me] = send_off;
    g->send_bytes[me] = send_bytes;
    if (!g->barrier()) return fail(ctx, PCUDA_ERR_NCCL, "local communicator: a rank did not reach the collective");
    for (int k = 0; k < g->world; ++k) {
        const int p = (me + k) % g->world;
        const size_t bytes = g->send_bytes[p][me];
        if (bytes != recv_bytes[p]) {
            if (status == PCUDA_OK)
                status = fail(ctx, PCUDA_ERR_NCCL, "local communicator: rank %d sends %zu bytes to rank %d, which expects %zu",
                              p, bytes, me, recv_bytes[p]);
            continue;
        }
        const char *src = g->send[p] + g->send_off[p][me];
        char *dst = static_cast<char *>(d_recv) + recv_off[p];
        if (!bytes || src == dst) continue;  // nothing to move / in-place own share
        if (p != me) check(cudaStreamWaitEvent(ctx->stream, g->ready[p], 0));
        check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    check(cudaEventRecord(g->done[me], ctx->stream));
    // every rank has queued its copies (the published offset tables may go out of scope now)
    if (!g->barrier()) return fail(ctx, PCUDA_ERR_NCCL, "local communicator: a rank did not reach the collective");
    for (int p = 0; p < g->world; ++p)
        if (p != me) check(cudaStreamWaitEvent(ctx->stream, g->done[p], 0));
    return status;
}

static int local_allgather(pcuda_ctx *ctx, const void *d_send, void *d_recv, size_t bytes) {
    size_t zero[LOCAL_MAX], cnt[LOCAL_MAX], roff[LOCAL_MAX];
    for (int p = 0; p < ctx->nccl->world; ++p) {
        zero[p] = 0;
        cnt[p] = bytes;
        roff[p] = (size_t)p * bytes;
    }
    return local_alltoallv(ctx, d_send, zero, cnt, d_recv, roff, cnt);
}

// Several exchanges as one NCCL group (the calls between begin and end are fused into one launch); nothing
// to do for the in-process communicator.
int nccl_group_begin(pcuda_ctx *ctx) {
    Nccl *n = ctx->nccl;
    if (!n || n->local || !n->comm || !n->GroupStart) return PCUDA_OK;
    ncclResult_t rc = n->GroupStart();
    return rc ? nccl_fail(ctx, "ncclGroupStart", rc) : PCUDA_OK;
}

int nccl_group_end(pcuda_ctx *ctx) {
    Nccl *n = ctx->nccl;
    if (!n || n->local || !n->comm || !n->GroupEnd) return PCUDA_OK;
    ncclResult_t rc = n->GroupEnd();
    return rc ? nccl_fail(ctx, "ncclGroupEnd", rc) : PCUDA_OK;
}

// Variable all-to-all on the context stream: rank o gets send_bytes[o] bytes from d_send +
// send_off[o] and delivers recv_bytes[o] bytes to d_recv + recv_off[o]; one grouped batch of
// ncclSend / ncclRecv.  The own share is a device-to-device copy.  Zero-byte pairs are skipped
// (both sides know the whole matrix, so they skip the same pairs).
int nccl_alltoallv(pcuda_ctx *ctx, const void *d_send, const size_t *send_off, const size_t *send_bytes,
                   void *d_recv, const size_t *recv_off, const size_t *recv_bytes) {
    if (!nccl_has_p2p(ctx)) return fail(ctx, PCUDA_ERR_NCCL, "ncclSend / ncclRecv are not available");
    Nccl *n = ctx->nccl;
    if (n->local) return local_alltoallv(ctx, d_send, send_off, send_bytes, d_recv, recv_off, recv_bytes);
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

int pcuda_comm_init_local(pcuda_ctx *const *ctxs, int world_size) {
    if (!ctxs || world_size < 1 || world_size > LOCAL_MAX)
        return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "local communicator: 1..%d contexts", LOCAL_MAX);
    for (int r = 0; r < world_size; ++r)
        if (!ctxs[r]) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "local communicator: NULL context");
    LocalGroup *g = new LocalGroup();
    g->world = world_size;
    g->attached = world_size;
    for (int r = 0; r < world_size; ++r) {
        pcuda_ctx *ctx = ctxs[r];
        nccl_free(ctx);
        DeviceGuard guard(ctx->device);
        Nccl *n = new Nccl();
        n->local = g;
        n->world = world_size;
        n->rank = r;
        ctx->nccl = n;
        cudaError_t e = cudaEventCreateWithFlags(&g->ready[r], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->done[r], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail(ctx, PCUDA_ERR_CUDA, "local communicator: %s", cudaGetErrorString(e));
    }
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
    if (ctx->nccl && ctx->nccl->local) {
        DeviceGuard guard(ctx->device);
        return local_allgather(ctx, d_send, d_recv, bytes_per_rank);
    }
    if (!ctx->nccl || !ctx->nccl->comm)
        return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "pcuda_comm_init has not been called");
    DeviceGuard guard(ctx->device);
    ncclResult_t r = ctx->nccl->AllGather(d_send, d_recv, bytes_per_rank, ncclInt8, ctx->nccl->comm,
                                          ctx->stream);
    if (r) return nccl_fail(ctx, "ncclAllGather", r);
    return PCUDA_OK;
}

}  // extern "C"
