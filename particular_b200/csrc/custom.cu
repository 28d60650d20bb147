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

// One thread per affected element; affecting elements staged through shared memory one tile at a
// time as raw 32-bit words (any trivially copyable struct whose size is a multiple of 4 bytes).
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
constexpr int PCUDA_BLOCK = 128;
constexpr int PCUDA_TILE = (sizeof(Affecting) * 128 <= 16384) ? 128 : 32;

__device__ inline void pcuda_compute_fwd(const Affected &p1, const Affecting &p2, Interaction &out) {
    compute(p1, p2, out);
}

extern "C" __global__ void __launch_bounds__(PCUDA_BLOCK)
pcuda_custom_brute_force(const Affected *__restrict__ affected, unsigned n_affected,
                         const Affecting *__restrict__ affecting, unsigned n_affecting,
                         Interaction *__restrict__ interactions) {
    __shared__ __align__(16) unsigned tile_words[PCUDA_TILE * (sizeof(Affecting) / 4)];
    const Affecting *tile = reinterpret_cast<const Affecting *>(tile_words);
    const unsigned i = blockIdx.x * PCUDA_BLOCK + threadIdx.x;
    const bool live = i < n_affected;
    Affected p1;
    if (live) p1 = affected[i];
    Interaction out = Interaction();
    constexpr unsigned WORDS = sizeof(Affecting) / 4;
    for (unsigned first = 0; first < n_affecting; first += PCUDA_TILE) {
        const unsigned cnt = min((unsigned)PCUDA_TILE, n_affecting - first);
        const unsigned *src = reinterpret_cast<const unsigned *>(affecting + first);
        __syncthreads();
        for (unsigned w = threadIdx.x; w < cnt * WORDS; w += PCUDA_BLOCK) tile_words[w] = src[w];
        __syncthreads();
        if (live)
            for (unsigned j = 0; j < cnt; ++j) pcuda_compute_fwd(p1, tile[j], out);
    }
    if (live) interactions[i] = out;
}

extern "C" __global__ void pcuda_custom_sizes(unsigned *out) {
    out[0] = sizeof(Affected);
    out[1] = sizeof(Affecting);
    out[2] = sizeof(Interaction);
    out[3] = sizeof(Push);
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
    uint32_t sizes[4] = {0, 0, 0, 0};  // Affected, Affecting, Interaction, Push
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
    if (ctx->d_misc.ensure(16) != cudaSuccess) return bail(fail(ctx, PCUDA_ERR_OUT_OF_MEMORY, "scratch"));
    void *d_sizes = ctx->d_misc.p;
    void *args[] = {&d_sizes};
    r = custom::g_api.LaunchKernel(sizes_fn, 1, 1, 1, 1, 1, 1, 0, (CUstream)ctx->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return bail(custom::cu_fail(ctx, "cuLaunchKernel(sizes)", r));
    if (cudaMemcpyAsync(it->sizes, d_sizes, 16, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return bail(fail(ctx, PCUDA_ERR_CUDA, "reading the struct sizes failed"));
    *out = it;
    return PCUDA_OK;
}

int pcuda_interaction_sizes(const pcuda_interaction *it, uint32_t sizes[4]) {
    if (!it || !sizes) return PCUDA_ERR_INVALID_ARGUMENT;
    memcpy(sizes, it->sizes, sizeof it->sizes);
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
    PCUDA_CUDA_TRY(ctx, it->d_affecting.ensure(b_bytes ? b_bytes : 4));
    PCUDA_CUDA_TRY(ctx, it->d_out.ensure(o_bytes));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(it->d_affected.p, affected, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (b_bytes)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(it->d_affecting.p, affecting, b_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (push_bytes)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(reinterpret_cast<void *>(it->push_ptr), push, push_bytes,
                                            cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    phase_begin(ctx, PH_COMPUTE);
    void *d_a = it->d_affected.p, *d_b = it->d_affecting.p, *d_o = it->d_out.p;
    unsigned na = (unsigned)n_affected, nb = (unsigned)n_affecting;
    void *args[] = {&d_a, &na, &d_b, &nb, &d_o};
    CUresult r = custom::g_api.LaunchKernel(it->kernel, (na + 127) / 128, 1, 1, 128, 1, 1, 0,
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
