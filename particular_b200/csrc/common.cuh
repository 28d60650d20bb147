// common.cuh — internal declarations shared by the translation units of libparticular_cuda.so.
// Not part of the C ABI (that is include/particular_cuda.h).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/particular_cuda.h"

namespace pcuda {

// Grow-only device buffer (the reference re-creates its wgpu buffers on any size change,
// gpu/resources.rs:26-34; we only ever grow).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return static_cast<T *>(p);
    }
};

struct Nccl;  // comm.cu

enum Phase { PH_UPLOAD = 0, PH_COMM, PH_BUILD, PH_COMPUTE, PH_DOWNLOAD, PH_COMM2, PH_COMM3, PH_COUNT };  // COMM2, COMM3: further exchanges of the call (added to comm_ms)

}  // namespace pcuda

struct pcuda_forest;  // barneshut.cu: buffers of the key-range-partitioned multi-GPU build

struct pcuda_ctx {
    int device = 0;
    int sm_count = 0;
    int sm_clock_khz = 0;
    size_t smem_optin = 0;
    char name[128] = {0};
    uint32_t leaf_size = 16;
    uint32_t order = 1;         // Barnes-Hut expansion order (pcuda_config.expansion_order)
    bool phase_timings = true;  // PCUDA_FLAG_NO_PHASE_TIMINGS clears it
    bool exact_checked = false;  // PCUDA_FLAG_EXACT_CHECKED
    cudaStream_t stream = nullptr;
    // copy engines' streams + events of the chunked host path (created on first use, bruteforce.cu)
    cudaStream_t stream_h2d = nullptr, stream_d2h = nullptr;
    cudaEvent_t ev_chunk_up[2] = {nullptr, nullptr}, ev_chunk_done[2] = {nullptr, nullptr},
                ev_chunk_free[2] = {nullptr, nullptr}, ev_d2h_end = nullptr;
    std::string err;

    // phase timing: start/stop event per phase, recorded lazily
    cudaEvent_t ev0[pcuda::PH_COUNT] = {}, ev1[pcuda::PH_COUNT] = {};
    bool ev_used[pcuda::PH_COUNT] = {};
    pcuda_timings timings = {};
    uint32_t launches = 0;

    // brute force scratch
    pcuda::DevBuf d_affected, d_affecting, d_out, d_partial, d_packed_src, d_packed_tgt, d_massmax,
        d_tile_done;
    // Barnes-Hut scratch (barneshut.cu owns the layout)
    pcuda::DevBuf d_stack, d_counters, d_tgt_keys, d_tgt_keys_alt, d_tgt_perm, d_tgt_perm_alt,
        d_tgt_sorted, d_cub_tmp, d_misc;
    uint64_t last_counters[5] = {0, 0, 0, 0, 0};
    pcuda_tree *call_tree = nullptr;  // tree reused by the one-shot Barnes-Hut entry points
    pcuda_forest *forest = nullptr;   // partitioned multi-GPU build (PCUDA_FLAG_BH_PARTITIONED_BUILD)
    int bh_build = 0;                 // multi-GPU tree build: 0 = automatic, 1 = partitioned, 2 = replicated, 3 = LET

    pcuda::Nccl *nccl = nullptr;
    int live_sims = 0;  // pcuda_sim objects created on this context (sim.cu)
};

namespace pcuda {

int fail(pcuda_ctx *ctx, int status, const char *fmt, ...);
void set_thread_error(const char *msg);

void phase_begin(pcuda_ctx *ctx, Phase p);
void phase_end(pcuda_ctx *ctx, Phase p);
void timings_reset(pcuda_ctx *ctx);
// Synchronises the stream and folds event times into ctx->timings.
int timings_collect(pcuda_ctx *ctx);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Internal enqueue-only entry points shared with sim.cu (device pointers, context stream, no
// synchronisation, no timing).  Target rows have `tgt_stride` scalars (positions first); source
// rows are {position, mu} = dim + 1 scalars.
int bf_enqueue_f32(pcuda_ctx *ctx, int dim, const float *d_tgt, int tgt_stride, size_t na,
                   const float *d_src, size_t nb, float softening, int checked, float *d_out);
int bf_enqueue_f64(pcuda_ctx *ctx, int dim, const double *d_tgt, int tgt_stride, size_t na,
                   const double *d_src, size_t nb, double softening, int checked, double *d_out);
// tgt_stride == 0: the targets are the sources themselves (the `&[P]` storage).
int bh_enqueue_f32(pcuda_ctx *ctx, int dim, const float *d_tgt, int tgt_stride, size_t na,
                   const float *d_src, size_t nb, float theta, float softening, float *d_out);
int bh_enqueue_f64(pcuda_ctx *ctx, int dim, const double *d_tgt, int tgt_stride, size_t na,
                   const double *d_src, size_t nb, double theta, double softening, double *d_out);

void tree_free(pcuda_ctx *ctx, pcuda_tree *t);
void forest_free(pcuda_ctx *ctx);
int bh_debug_set(const char *key, int value);  // barneshut.cu tuning hooks
void nccl_free(pcuda_ctx *ctx);
// comm.cu: world size / rank of the context's communicator (1 / 0 when none was initialised).
void nccl_world(const pcuda_ctx *ctx, int *world, int *rank);
bool nccl_has_p2p(const pcuda_ctx *ctx);
void nccl_poison(pcuda_ctx *ctx);  // after a step failed between collectives: abort, fail fast from now on
bool nccl_poisoned(const pcuda_ctx *ctx);
int nccl_group_begin(pcuda_ctx *ctx);  // the exchanges up to nccl_group_end() become one NCCL group
int nccl_group_end(pcuda_ctx *ctx);
int nccl_alltoallv(pcuda_ctx *ctx, const void *d_send, const size_t *send_off, const size_t *send_bytes,
                   void *d_recv, const size_t *recv_off, const size_t *recv_bytes);

}  // namespace pcuda

#define PCUDA_CUDA_TRY(ctx, expr)                                                              \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return pcuda::fail((ctx),                                                          \
                               _e == cudaErrorMemoryAllocation ? PCUDA_ERR_OUT_OF_MEMORY       \
                                                               : PCUDA_ERR_CUDA,               \
                               "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,              \
                               cudaGetErrorString(_e));                                        \
    } while (0)

#define PCUDA_TRY(expr)                 \
    do {                                \
        int _s = (expr);                \
        if (_s != PCUDA_OK) return _s;  \
    } while (0)
