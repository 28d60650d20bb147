// custom.cu — user-defined pair interactions on the device.
//
// The reference's GPU operator is generic over the interaction: an `InteractionShader<P1, P2>`
// supplies WGSL source that defines the types `Affected`, `Affecting`, `Interaction` and a function
// `compute(p1, p2, out)`, the byte sizes of the three types, writers/readers for the buffers and
// optional push constants (particular/src/gpu/mod.rs:40-82); the operator pastes it into a
// brute-force template (gpu/bruteforce.wgsl: one invocation per affected particle, a loop over all
// affecting particles, `var out = Interaction()`), compiles it at run time and dispatches it
// (gpu/resources.rs:126-233).  This is the CUDA counterpart: the user supplies CUDA C++ defining
//
//     struct Affected { ... };  struct Affecting { ... };  struct Interaction { ... };
//     struct Push { ... };                      // "push constants"; may be empty
//     __device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out);
//
// (`push` is visible to `compute` as a __constant__ object), the library pastes it into the
// brute-force template below, compiles it with NVRTC for sm_100a, loads the cubin and launches it.
// NVRTC and the driver API are dlopen()ed on first use, so the library itself links against
// neither.  Gravity does not go through here: Acceleration / AccelerationSoftened have the
// hand-tuned kernels of bruteforce.cu.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace pcuda {
namespace custom {

// The template is the skeleton of the hand-written pair kernel (bruteforce.cu), made generic: the
// affecting elements stream through a 4-stage shared-memory ring filled by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx; full / empty barrier pairs, one elected producer thread), every
// thread owns PCUDA_TP affected elements and their accumulators in registers, so one shared-memory read
// of an affecting element feeds PCUDA_TP calls of compute().  The reference's template is one invocation
// per affected element with a workgroup-shared tile and two barriers per tile
// (gpu/bruteforce_shared.wgsl:1-31).  Every accumulator still folds the affecting elements in slice
// order from `Interaction()` — the order of sequential::BruteForce (sequential.rs:181-194) — which is
// what makes the gravity term written this way bit-identical to the CPU fold; that is also why the
// sources are not split over CTAs here as the gravity kernel does.
static const char *kPrologue = R"PCUDA(
#define PCUDA_CUSTOM 1
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
struct Push;                      // defined by the interaction source, before compute()
extern __constant__ Push push;    // the "push constants" (gpu/mod.rs:72-81), set per call
)PCUDA";

static const char *kTemplate = R"PCUDA(
__constant__ Push push;
static_assert(sizeof(Affected) % 4 == 0 && sizeof(Affecting) % 4 == 0 && sizeof(Interaction) % 4 == 0,
              "Affected / Affecting / Interaction sizes must be multiples of 4 bytes");
static_assert(sizeof(Affecting) <= 256, "Affecting must not exceed 256 bytes");
constexpr int PCUDA_BLOCK = 128;
constexpr int PCUDA_STAGES = 4;
constexpr int PCUDA_PREFETCH = 2;
// elements per stage: a multiple of 4 (so that a stage is a multiple of 16 bytes, the TMA granule)
constexpr int PCUDA_TILE = sizeof(Affecting) <= 64 ? 128 : 32;
// affected elements per thread: two while their state plausibly stays in registers
constexpr int PCUDA_TP = (sizeof(Affected) + sizeof(Interaction) <= 64) ? 2 : 1;

__device__ inline void pcuda_compute_fwd(const Affected &p1, const Affecting &p2, Interaction &out) {
    compute(p1, p2, out);
}

__device__ __forceinline__ unsigned pcuda_smem(const void *p) {
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void pcuda_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pcuda_smem(bar)), "r"(count));
}
__device__ __forceinline__ void pcuda_mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pcuda_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pcuda_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pcuda_smem(bar)) : "memory");
}
__device__ __forceinline__ void pcuda_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PCUDA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PCUDA_DONE;\n"
        "bra PCUDA_WAIT;\n"
        "PCUDA_DONE:\n"
        "}\n" ::"r"(pcuda_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void pcuda_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pcuda_smem(dst)), "l"(src), "r"(bytes), "r"(pcuda_smem(bar)) : "memory");
}

// `affecting` is padded by the host to a whole number of tiles (the padding is never passed to compute()).
extern "C" __global__ void __launch_bounds__(PCUDA_BLOCK)
pcuda_custom_brute_force(const Affected *__restrict__ affected, unsigned n_affected,
                         const Affecting *__restrict__ affecting, unsigned n_affecting,
                         Interaction *__restrict__ interactions) {
    __shared__ __align__(128) unsigned char tile_bytes[PCUDA_STAGES][PCUDA_TILE * sizeof(Affecting)];
    __shared__ __align__(8) unsigned long long full_bar[PCUDA_STAGES], empty_bar[PCUDA_STAGES];
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const unsigned base = blockIdx.x * (PCUDA_BLOCK * PCUDA_TP) + tid;
    Affected p1[PCUDA_TP];
    Interaction out[PCUDA_TP];
#pragma unroll
    for (int p = 0; p < PCUDA_TP; ++p) {
        const unsigned i = base + p * PCUDA_BLOCK;
        p1[p] = affected[i < n_affected ? i : n_affected - 1];
        out[p] = Interaction();
    }
    const unsigned ntiles = (n_affecting + PCUDA_TILE - 1) / PCUDA_TILE;
    constexpr unsigned STAGE_BYTES = PCUDA_TILE * sizeof(Affecting);
    auto issue = [&](unsigned t) {  // the elected producer: one bulk copy per tile
        const unsigned st = t % PCUDA_STAGES;
        pcuda_mbar_expect(&full_bar[st], STAGE_BYTES);
        pcuda_bulk_g2s(tile_bytes[st], reinterpret_cast<const unsigned char *>(affecting) + (size_t)t * STAGE_BYTES,
                       STAGE_BYTES, &full_bar[st]);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < PCUDA_STAGES; ++s) {
            pcuda_mbar_init(&full_bar[s], 1);
            pcuda_mbar_init(&empty_bar[s], PCUDA_BLOCK / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (unsigned t = 0; t < PCUDA_PREFETCH && t < ntiles; ++t) issue(t);
    for (unsigned t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + PCUDA_PREFETCH < ntiles) {
            const unsigned tn = t + PCUDA_PREFETCH;
            if (tn >= PCUDA_STAGES) pcuda_mbar_wait(&empty_bar[tn % PCUDA_STAGES], ((tn / PCUDA_STAGES) - 1) & 1);
            issue(tn);
        }
        const unsigned st = t % PCUDA_STAGES;
        pcuda_mbar_wait(&full_bar[st], (t / PCUDA_STAGES) & 1);
        const Affecting *tile = reinterpret_cast<const Affecting *>(tile_bytes[st]);
        const unsigned cnt = min((unsigned)PCUDA_TILE, n_affecting - t * PCUDA_TILE);
        for (unsigned j = 0; j < cnt; ++j) {
            const Affecting p2 = tile[j];  // one (broadcast) shared-memory read feeds PCUDA_TP interactions
#pragma unroll
            for (int p = 0; p < PCUDA_TP; ++p) pcuda_compute_fwd(p1[p], p2, out[p]);
        }
        __syncwarp();
        if (lane == 0) pcuda_mbar_arrive(&empty_bar[st]);
    }
#pragma unroll
    for (int p = 0; p < PCUDA_TP; ++p) {
        const unsigned i = base + p * PCUDA_BLOCK;
        if (i < n_affected) interactions[i] = out[p];
    }
}

extern "C" __global__ void pcuda_custom_sizes(unsigned *out) {
    out[0] = sizeof(Affected);
    out[1] = sizeof(Affecting);
    out[2] = sizeof(Interaction);
    out[3] = sizeof(Push);
    out[4] = PCUDA_TILE;
    out[5] = PCUDA_TP;
}
)PCUDA";

struct Api {
    void *nvrtc = nullptr, *cuda = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*ModuleGetGlobal)(CUdeviceptr *, size_t *, CUmodule, const char *) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void **, void **) = nullptr;
    CUresult (*CuGetErrorString)(CUresult, const char **) = nullptr;
};

static Api g_api;
static std::mutex g_api_mutex;

template <class F>
static bool sym(void *lib, const char *name, F &out) {
    out = reinterpret_cast<F>(dlsym(lib, name));
    return out != nullptr;
}

static int load_nvrtc(pcuda_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (g_api.nvrtc) return PCUDA_OK;
    void *lib = nullptr;
    for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12",
                             "/usr/local/cuda/lib64/libnvrtc.so"}) {
        lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "libnvrtc.so not found: %s", dlerror());
    Api &a = g_api;
    const bool ok = sym(lib, "nvrtcCreateProgram", a.CreateProgram) & sym(lib, "nvrtcCompileProgram", a.CompileProgram) &
                    sym(lib, "nvrtcGetCUBINSize", a.GetCUBINSize) & sym(lib, "nvrtcGetCUBIN", a.GetCUBIN) &
                    sym(lib, "nvrtcGetProgramLogSize", a.GetProgramLogSize) &
                    sym(lib, "nvrtcGetProgramLog", a.GetProgramLog) &
                    sym(lib, "nvrtcDestroyProgram", a.DestroyProgram) &
                    sym(lib, "nvrtcGetErrorString", a.GetErrorString);
    if (!ok) {
        dlclose(lib);
        return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "libnvrtc.so lacks a required symbol");
    }
    a.nvrtc = lib;
    return PCUDA_OK;
}

static int load_driver(pcuda_ctx *ctx) {
    std::lock_guard<std::mutex> lock(g_api_mutex);
    if (g_api.cuda) return PCUDA_OK;
    void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!lib) return fail(ctx, PCUDA_ERR_NO_DEVICE, "libcuda.so.1 not found: %s", dlerror());
    Api &a = g_api;
    const bool ok = sym(lib, "cuModuleLoadData", a.ModuleLoadData) & sym(lib, "cuModuleUnload", a.ModuleUnload) &
                    sym(lib, "cuModuleGetFunction", a.ModuleGetFunction) &
                    sym(lib, "cuModuleGetGlobal_v2", a.ModuleGetGlobal) &
                    sym(lib, "cuLaunchKernel", a.LaunchKernel) & sym(lib, "cuGetErrorString", a.CuGetErrorString);
    if (!ok) {
        dlclose(lib);
        return fail(ctx, PCUDA_ERR_NOT_INITIALISED, "libcuda.so.1 lacks a required symbol");
    }
    a.cuda = lib;
    return PCUDA_OK;
}

static int cu_fail(pcuda_ctx *ctx, const char *what, CUresult r) {
    const char *msg = nullptr;
    if (g_api.CuGetErrorString) g_api.CuGetErrorString(r, &msg);
    return fail(ctx, PCUDA_ERR_CUDA, "%s failed: %s (%d)", what, msg ? msg : "?", (int)r);
}

// Pastes the user source into the template and compiles it to an sm_100a cubin.  On a compile error
// the NVRTC log becomes the error message (the reference panics with naga's validation error).
static int compile(pcuda_ctx *ctx, const char *source, std::vector<char> &cubin, std::string *log_out) {
    PCUDA_TRY(load_nvrtc(ctx));
    std::string full = std::string(kPrologue) + "\n#line 1 \"interaction.cu\"\n" + source +
                       "\n#line 1 \"pcuda_template.cu\"\n" + kTemplate;
    nvrtcProgram prog = nullptr;
    nvrtcResult r = g_api.CreateProgram(&prog, full.c_str(), "pcuda_custom.cu", 0, nullptr, nullptr);
    if (r != NVRTC_SUCCESS) return fail(ctx, PCUDA_ERR_CUDA, "nvrtcCreateProgram: %s", g_api.GetErrorString(r));
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "--fmad=false"};
    r = g_api.CompileProgram(prog, 4, opts);
    std::string log;
    size_t log_size = 0;
    if (g_api.GetProgramLogSize(prog, &log_size) == NVRTC_SUCCESS && log_size > 1) {
        log.resize(log_size);
        g_api.GetProgramLog(prog, &log[0]);
        while (!log.empty() && (log.back() == '\0' || log.back() == '\n')) log.pop_back();
    }
    if (log_out) *log_out = log;
    if (r != NVRTC_SUCCESS) {
        g_api.DestroyProgram(&prog);
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "interaction source does not compile (%s):\n%s",
                    g_api.GetErrorString(r), log.c_str());
    }
    size_t size = 0;
    r = g_api.GetCUBINSize(prog, &size);
    if (r == NVRTC_SUCCESS) {
        cubin.resize(size);
        r = g_api.GetCUBIN(prog, cubin.data());
    }
    g_api.DestroyProgram(&prog);
    if (r != NVRTC_SUCCESS) return fail(ctx, PCUDA_ERR_CUDA, "nvrtcGetCUBIN: %s", g_api.GetErrorString(r));
    return PCUDA_OK;
}

}  // namespace custom
}  // namespace pcuda

struct pcuda_interaction {
    CUmodule module = nullptr;
    CUfunction kernel = nullptr;
    CUdeviceptr push_ptr = 0;
    uint32_t sizes[6] = {0, 0, 0, 0, 0, 0};  // Affected, Affecting, Interaction, Push; tile, affected per thread
    pcuda::DevBuf d_affected, d_affecting, d_out;
};

using namespace pcuda;

extern "C" {

int pcuda_interaction_check(const char *source, char *log, size_t log_len) {
    if (!source) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "source is NULL");
    std::vector<char> cubin;
    std::string text;
    int s = custom::compile(nullptr, source, cubin, &text);
    if (log && log_len) {
        strncpy(log, text.c_str(), log_len - 1);
        log[log_len - 1] = 0;
    }
    return s;
}

int pcuda_interaction_create(pcuda_ctx *ctx, const char *source, pcuda_interaction **out) {
    if (!ctx || !source || !out) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    std::vector<char> cubin;
    PCUDA_TRY(custom::compile(ctx, source, cubin, nullptr));
    PCUDA_TRY(custom::load_driver(ctx));
    PCUDA_CUDA_TRY(ctx, cudaFree(nullptr));  // makes the runtime's primary context current for the driver calls
    pcuda_interaction *it = new pcuda_interaction();
    auto bail = [&](int status) {
        pcuda_interaction_destroy(ctx, it);
        return status;
    };
    CUresult r = custom::g_api.ModuleLoadData(&it->module, cubin.data());
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuModuleLoadData", r));
    r = custom::g_api.ModuleGetFunction(&it->kernel, it->module, "pcuda_custom_brute_force");
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuModuleGetFunction", r));
    size_t push_size = 0;
    r = custom::g_api.ModuleGetGlobal(&it->push_ptr, &push_size, it->module, "push");
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuModuleGetGlobal(push)", r));
    // the struct sizes as the device compiler sees them
    CUfunction sizes_fn = nullptr;
    r = custom::g_api.ModuleGetFunction(&sizes_fn, it->module, "pcuda_custom_sizes");
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuModuleGetFunction(sizes)", r));
    if (ctx->d_misc.ensure(32) != cudaSuccess) return bail(fail(ctx, PCUDA_ERR_OUT_OF_MEMORY, "scratch"));
    void *d_sizes = ctx->d_misc.p;
    void *args[] = {&d_sizes};
    r = custom::g_api.LaunchKernel(sizes_fn, 1, 1, 1, 1, 1, 1, 0, (CUstream)ctx->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuLaunchKernel(sizes)", r));
    if (cudaMemcpyAsync(it->sizes, d_sizes, 24, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, PCUDA_ERR_CUDA, "reading the struct sizes failed"));
    *out = it;
    return PCUDA_OK;
}

int pcuda_interaction_sizes(const pcuda_interaction *it, uint32_t sizes[4]) {
    if (!it || !sizes) return PCUDA_ERR_INVALID_ARGUMENT;
    memcpy(sizes, it->sizes, 4 * sizeof(uint32_t));
    return PCUDA_OK;
}

int pcuda_interaction_brute_force(pcuda_ctx *ctx, pcuda_interaction *it, const void *affected,
                                  size_t n_affected, const void *affecting, size_t n_affecting,
                                  const void *push, size_t push_bytes, void *out) {
    if (!ctx || !it) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    if ((n_affected && (!affected || !out)) || (n_affecting && !affecting))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (n_affected > 0x7fffffffull || n_affecting > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    if (push_bytes > it->sizes[3] || (push_bytes && !push))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "push constants: %zu bytes given, struct Push has %u",
                    push_bytes, it->sizes[3]);
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (n_affected == 0) return PCUDA_OK;
    const size_t a_bytes = n_affected * it->sizes[0], b_bytes = n_affecting * it->sizes[1],
                 o_bytes = n_affected * it->sizes[2];
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, it->d_affected.ensure(a_bytes));
    // the kernel copies whole tiles (TMA bulk copies): pad the affecting buffer to a tile boundary
    const size_t tile_bytes = (size_t)it->sizes[4] * it->sizes[1];
    const size_t b_padded = (b_bytes + tile_bytes - 1) / tile_bytes * tile_bytes;
    PCUDA_CUDA_TRY(ctx, it->d_affecting.ensure(b_padded ? b_padded : 16));
    PCUDA_CUDA_TRY(ctx, it->d_out.ensure(o_bytes));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(it->d_affected.p, affected, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (b_bytes) {
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(it->d_affecting.p, affecting, b_bytes, cudaMemcpyHostToDevice, ctx->stream));
        if (b_padded > b_bytes)
            PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(static_cast<char *>(it->d_affecting.p) + b_bytes, 0, b_padded - b_bytes,
                                                ctx->stream));
    }
    if (push_bytes)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<void *>(it->push_ptr), push, push_bytes,
                                            cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    phase_begin(ctx, PH_COMPUTE);
    void *d_a = it->d_affected.p, *d_b = it->d_affecting.p, *d_o = it->d_out.p;
    unsigned na = (unsigned)n_affected, nb = (unsigned)n_affecting;
    void *args[] = {&d_a, &na, &d_b, &nb, &d_o};
    const unsigned per_cta = 128 * it->sizes[5];
    CUresult r = custom::g_api.LaunchKernel(it->kernel, (na + per_cta - 1) / per_cta, 1, 1, 128, 1, 1, 0,
                                            (CUstream)ctx->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return custom::cu_fail(ctx, "cuLaunchKernel", r);
    ctx->launches++;
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out, it->d_out.p, o_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

void pcuda_interaction_destroy(pcuda_ctx *ctx, pcuda_interaction *it) {
    if (!it) return;
    if (ctx) {
        DeviceGuard guard(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (it->module && custom::g_api.ModuleUnload) custom::g_api.ModuleUnload(it->module);
        DevBuf *bufs[] = {&it->d_affected, &it->d_affecting, &it->d_out};
        for (DevBuf *b : bufs) b->release();
    }
    delete it;
}

}  // extern "C"
