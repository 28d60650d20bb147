// context.cu — context lifetime, error reporting, phase timing.  C ABI: include/particular_cuda.h.
#include <cstring>

#include "common.cuh"

namespace pcuda {

static thread_local std::string g_thread_error;

void set_thread_error(const char *msg) { g_thread_error = msg ? msg : ""; }

int fail(pcuda_ctx *ctx, int status, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_thread_error = buf;
    // A failed launch leaves a sticky-free error in the runtime's per-thread slot: clear it.
    cudaGetLastError();
    return status;
}

void timings_reset(pcuda_ctx *ctx) {
    for (int i = 0; i < PH_COUNT; ++i) ctx->ev_used[i] = false;
    ctx->timings = pcuda_timings{};
    ctx->launches = 0;
}

void phase_begin(pcuda_ctx *ctx, Phase p) {
    if (ctx->phase_timings && !ctx->ev_used[p]) cudaEventRecord(ctx->ev0[p], ctx->stream);
}

void phase_end(pcuda_ctx *ctx, Phase p) {
    if (!ctx->phase_timings) return;
    cudaEventRecord(ctx->ev1[p], ctx->stream);
    ctx->ev_used[p] = true;
}

int timings_collect(pcuda_ctx *ctx) {
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    float *dst[PH_COUNT] = {&ctx->timings.upload_ms, &ctx->timings.comm_ms, &ctx->timings.build_ms,
                            &ctx->timings.compute_ms, &ctx->timings.download_ms, nullptr, nullptr};
    ctx->timings = pcuda_timings{};
    for (int i = 0; i < PH_COUNT; ++i) {
        if (!ctx->ev_used[i]) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev0[i], ctx->ev1[i]) != cudaSuccess) continue;
        *(dst[i] ? dst[i] : &ctx->timings.comm_ms) += ms;
    }
    ctx->timings.kernel_launches = ctx->launches;
    return PCUDA_OK;
}

}  // namespace pcuda

using namespace pcuda;

extern "C" {

int pcuda_abi_version(void) { return PCUDA_ABI_VERSION; }

const char *pcuda_status_string(int status) {
    switch (status) {
        case PCUDA_OK: return "ok";
        case PCUDA_ERR_INVALID_ARGUMENT: return "invalid argument";
        case PCUDA_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
        case PCUDA_ERR_CUDA: return "CUDA runtime error";
        case PCUDA_ERR_OUT_OF_MEMORY: return "out of device memory";
        case PCUDA_ERR_NCCL: return "NCCL error";
        case PCUDA_ERR_TREE_OVERFLOW: return "Barnes-Hut traversal stack overflow";
        case PCUDA_ERR_NOT_INITIALISED: return "not initialised";
        default: return "unknown status";
    }
}

int pcuda_device_count(int *count) {
    if (!count) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, PCUDA_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return PCUDA_OK;
}

int pcuda_create(const pcuda_config *config, pcuda_ctx **out) {
    if (!out) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int dev = config ? config->device : 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, PCUDA_ERR_NO_DEVICE,
                    "no CUDA device (%s); this backend has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (dev < 0 || dev >= n)
        return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", dev, n);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess)
        return fail(nullptr, PCUDA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, PCUDA_ERR_NO_DEVICE,
                    "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev,
                    prop.name, prop.major, prop.minor);

    pcuda_ctx *ctx = new pcuda_ctx();
    ctx->device = dev;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    ctx->sm_clock_khz = khz;
    strncpy(ctx->name, prop.name, sizeof(ctx->name) - 1);
    if (config && config->leaf_size) ctx->leaf_size = config->leaf_size;
    if (config && config->expansion_order > 2) {
        delete ctx;
        return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "expansion_order must be 0, 1 or 2");
    }
    if (config && config->expansion_order == 2) ctx->order = 2;
    ctx->phase_timings = !(config && (config->flags & PCUDA_FLAG_NO_PHASE_TIMINGS));
    ctx->exact_checked = config && (config->flags & PCUDA_FLAG_EXACT_CHECKED);
    ctx->bh_build = !config ? 0
                    : (config->flags & PCUDA_FLAG_BH_LET_BUILD)         ? 3
                    : (config->flags & PCUDA_FLAG_BH_PARTITIONED_BUILD) ? 1
                    : (config->flags & PCUDA_FLAG_BH_REPLICATED_BUILD)  ? 2
                                                                        : 0;
    if (ctx->leaf_size > 32) ctx->leaf_size = 32;

    DeviceGuard guard(dev);
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < PH_COUNT && e == cudaSuccess; ++i) {
        e = cudaEventCreate(&ctx->ev0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1[i]);
    }
    if (e != cudaSuccess) {
        int s = fail(nullptr, PCUDA_ERR_CUDA, "context setup: %s", cudaGetErrorString(e));
        pcuda_destroy(ctx);
        return s;
    }
    *out = ctx;
    return PCUDA_OK;
}

void pcuda_destroy(pcuda_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    nccl_free(ctx);
    if (ctx->call_tree) tree_free(ctx, ctx->call_tree);
    forest_free(ctx);
    DevBuf *bufs[] = {&ctx->d_affected, &ctx->d_affecting, &ctx->d_out, &ctx->d_partial,
                      &ctx->d_packed_src, &ctx->d_packed_tgt, &ctx->d_massmax, &ctx->d_tile_done, &ctx->d_stack, &ctx->d_counters,
                      &ctx->d_tgt_keys, &ctx->d_tgt_keys_alt, &ctx->d_tgt_perm,
                      &ctx->d_tgt_perm_alt, &ctx->d_tgt_sorted, &ctx->d_cub_tmp, &ctx->d_misc};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < PH_COUNT; ++i) {
        if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]);
        if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_chunk_up[i]) cudaEventDestroy(ctx->ev_chunk_up[i]);
        if (ctx->ev_chunk_done[i]) cudaEventDestroy(ctx->ev_chunk_done[i]);
        if (ctx->ev_chunk_free[i]) cudaEventDestroy(ctx->ev_chunk_free[i]);
    }
    if (ctx->ev_d2h_end) cudaEventDestroy(ctx->ev_d2h_end);
    if (ctx->stream_h2d) cudaStreamDestroy(ctx->stream_h2d);
    if (ctx->stream_d2h) cudaStreamDestroy(ctx->stream_d2h);

    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *pcuda_last_error(const pcuda_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_thread_error.c_str();
}

int pcuda_get_timings(const pcuda_ctx *ctx, pcuda_timings *out) {
    if (!ctx || !out) return PCUDA_ERR_INVALID_ARGUMENT;
    *out = ctx->timings;
    return PCUDA_OK;
}

void *pcuda_stream(pcuda_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int pcuda_sync(pcuda_ctx *ctx) {
    if (!ctx) return PCUDA_ERR_INVALID_ARGUMENT;
    DeviceGuard guard(ctx->device);
    return timings_collect(ctx);
}

int pcuda_device_info(const pcuda_ctx *ctx, int *sm_count, int *sm_clock_khz, char *name,
                      size_t name_len) {
    if (!ctx) return PCUDA_ERR_INVALID_ARGUMENT;
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_clock_khz) *sm_clock_khz = ctx->sm_clock_khz;
    if (name && name_len) {
        strncpy(name, ctx->name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return PCUDA_OK;
}

int pcuda_host_alloc(pcuda_ctx *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(ctx->device);
    PCUDA_CUDA_TRY(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PCUDA_OK;
}

int pcuda_host_free(pcuda_ctx *ctx, void *p) {
    if (!ctx) return PCUDA_ERR_INVALID_ARGUMENT;
    if (!p) return PCUDA_OK;
    DeviceGuard guard(ctx->device);
    PCUDA_CUDA_TRY(ctx, cudaFreeHost(p));
    return PCUDA_OK;
}

}  // extern "C"
