// sim.cu — device-resident stepping: particles, velocities and accelerations stay in HBM across
// steps; one step = accelerations (brute force or Barnes-Hut) + semi-implicit Euler.
//
// Every caller of the reference integrates the accelerations right after computing them
// (examples/simple/src/main.rs:45-59, examples/particle-toy/src/physics.rs:128-138, 166-176,
// the reference's own circular_orbit! test gravity/newtonian/mod.rs:318-331, and the benches add
// them to velocities, benches/benchmark.rs:81), and its wgpu operator pays an upload and a blocking
// read-back per call (gpu/resources.rs:37-39, 318-349).  Here nothing crosses PCIe per step.
//
// The integrator is the reference's, operation for operation and unfused:
//     velocity += acceleration * dt;  position += velocity * dt;
// so that, given the same accelerations, positions and velocities are bit-identical to the Rust loop.
//
// With PCUDA_SIM_AFFECTING_MASSIVE_ONLY the sources are the particles with mu != 0 in their original
// order — the `Reordered` storage (storage.rs:153-163, 219-229): affected = all particles in input
// order, affecting = the massive ones.  Masses never change, so the index list is built once and
// the compact source array is re-gathered from the current positions every step.
//
// Brute-force steps are replayed from a CUDA graph (8 steps per launch) once the scratch buffers
// have their final size: at the small particle counts of the reference's demos a step is
// launch-latency bound.
#include <cstring>

#include "common.cuh"

namespace pcuda {
namespace sim {

constexpr int GRAPH_STEPS = 8;

template <typename S>
__device__ __forceinline__ S mul_rn(S a, S b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename S>
__device__ __forceinline__ S add_rn(S a, S b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// v += a * dt; p += v * dt  (examples/simple/src/main.rs:55-58), one thread per scalar component.
template <typename S, int DIM>
__global__ void __launch_bounds__(256) integrate_kernel(S *__restrict__ particles, S *__restrict__ vel,
                                                        const S *__restrict__ acc, size_t n, S dt) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * DIM) return;
    const size_t i = t / DIM, c = t % DIM;
    const S v = add_rn(vel[t], mul_rn(acc[t], dt));
    vel[t] = v;
    S *p = particles + i * (DIM + 1) + c;
    *p = add_rn(*p, mul_rn(v, dt));
}

// Compact copy of the affecting rows: dst[j] = particles[idx[j]].
template <typename S, int ROW>
__global__ void __launch_bounds__(256) gather_rows(const S *__restrict__ particles,
                                                   const uint32_t *__restrict__ idx, size_t m,
                                                   S *__restrict__ dst) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * ROW) return;
    dst[t] = particles[(size_t)idx[t / ROW] * ROW + t % ROW];
}

}  // namespace sim
}  // namespace pcuda

struct pcuda_sim {
    pcuda_sim_config cfg = {};
    size_t n = 0, n_affecting = 0;
    bool subset = false;  // sources are a strict subset of the particles
    pcuda::DevBuf particles, vel, acc, idx, src;
    uint64_t steps_done = 0;
    bool warm = false;  // one eager step has sized every scratch buffer
    cudaGraphExec_t graph = nullptr;
    const void *graph_key[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // scratch pointers baked into the graph
    uint32_t launches_per_step = 0;
    size_t scalar_bytes() const { return cfg.scalar == PCUDA_F64 ? 8 : 4; }
};

namespace pcuda {
namespace sim {

static void drop_graph(pcuda_sim *s) {
    if (s->graph) cudaGraphExecDestroy(s->graph);
    s->graph = nullptr;
}

template <typename S, int DIM>
static int step_once_t(pcuda_ctx *ctx, pcuda_sim *s) {
    const size_t n = s->n;
    cudaStream_t st = ctx->stream;
    S *part = s->particles.as<S>();
    const S *src = part;
    size_t nb = n;
    if (s->subset) {
        nb = s->n_affecting;
        if (nb) {
            gather_rows<S, DIM + 1><<<(unsigned)((nb * (DIM + 1) + 255) / 256), 256, 0, st>>>(
                part, s->idx.as<uint32_t>(), nb, s->src.as<S>());
            PCUDA_CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        src = s->src.as<S>();
    }
    S *acc = s->acc.as<S>();
    if constexpr (sizeof(S) == 8) {
        if (s->cfg.algorithm == PCUDA_BARNES_HUT)
            PCUDA_TRY(bh_enqueue_f64(ctx, DIM, part, s->subset ? DIM + 1 : 0, n, src, nb, s->cfg.theta,
                                     s->cfg.softening, acc));
        else
            PCUDA_TRY(bf_enqueue_f64(ctx, DIM, part, DIM + 1, n, src, nb, (double)s->cfg.softening,
                                     s->cfg.checked, acc));
    } else if (s->cfg.algorithm == PCUDA_BARNES_HUT) {
        // targets == sources unless the sources are a subset
        PCUDA_TRY(bh_enqueue_f32(ctx, DIM, part, s->subset ? DIM + 1 : 0, n, src, nb,
                                 (float)s->cfg.theta, (float)s->cfg.softening, acc));
    } else {
        PCUDA_TRY(bf_enqueue_f32(ctx, DIM, part, DIM + 1, n, src, nb, (float)s->cfg.softening,
                                 s->cfg.checked, acc));
    }
    integrate_kernel<S, DIM><<<(unsigned)((n * DIM + 255) / 256), 256, 0, st>>>(
        part, s->vel.as<S>(), acc, n, (S)s->cfg.dt);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return PCUDA_OK;
}

static int step_once(pcuda_ctx *ctx, pcuda_sim *s) {
    if (s->cfg.scalar == PCUDA_F64)
        return s->cfg.dim == 3 ? step_once_t<double, 3>(ctx, s) : step_once_t<double, 2>(ctx, s);
    return s->cfg.dim == 3 ? step_once_t<float, 3>(ctx, s) : step_once_t<float, 2>(ctx, s);
}

static void scratch_key(const pcuda_ctx *ctx, const void *key[5]) {
    key[4] = ctx->d_tile_done.p;
    key[0] = ctx->d_partial.p;
    key[1] = ctx->d_massmax.p;
    key[2] = ctx->d_packed_src.p;
    key[3] = ctx->stream;
}

// Captures GRAPH_STEPS brute-force steps.  Returns false (and leaves the stream usable) when the
// capture cannot be made; the caller then runs eagerly.
static bool capture(pcuda_ctx *ctx, pcuda_sim *s) {
    drop_graph(s);
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const uint32_t before = ctx->launches;
    int status = PCUDA_OK;
    for (int k = 0; k < GRAPH_STEPS && status == PCUDA_OK; ++k) status = step_once(ctx, s);
    ctx->launches = before;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    if (status != PCUDA_OK || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return false;
    }
    e = cudaGraphInstantiate(&s->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        s->graph = nullptr;
        cudaGetLastError();
        return false;
    }
    scratch_key(ctx, s->graph_key);
    return true;
}

static int validate(pcuda_ctx *ctx, const pcuda_sim_config *c) {
    if (c->dim != 2 && c->dim != 3) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    if (c->scalar != PCUDA_F32 && c->scalar != PCUDA_F64)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "scalar must be PCUDA_F32 or PCUDA_F64");
    if (c->algorithm != PCUDA_BRUTE_FORCE && c->algorithm != PCUDA_BARNES_HUT)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "unknown algorithm %u", c->algorithm);
    if (c->algorithm == PCUDA_BARNES_HUT && !(c->theta >= 0.0))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "theta must be >= 0");
    return PCUDA_OK;
}

}  // namespace sim
}  // namespace pcuda

using namespace pcuda;

extern "C" {

int pcuda_sim_create(pcuda_ctx *ctx, const pcuda_sim_config *config, const void *particles,
                     const void *velocities, size_t n, pcuda_sim **out) {
    if (!ctx || !config || !out) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    if (n && !particles) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    PCUDA_TRY(sim::validate(ctx, config));
    DeviceGuard guard(ctx->device);
    pcuda_sim *s = new pcuda_sim();
    s->cfg = *config;
    s->n = n;
    const size_t sb = s->scalar_bytes(), dim = config->dim, row = (dim + 1) * sb;
    auto bail = [&](int status) {
        pcuda_sim_destroy(nullptr, s);
        return status;
    };
    // the affecting subset (mu != 0), found on the host once: masses never change
    std::vector<uint32_t> idx;
    if ((config->flags & PCUDA_SIM_AFFECTING_MASSIVE_ONLY) && n) {
        const char *p = static_cast<const char *>(particles);
        for (size_t i = 0; i < n; ++i) {
            bool massive;
            if (sb == 8) {
                double m;
                memcpy(&m, p + i * row + dim * sb, 8);
                massive = m != 0.0;
            } else {
                float m;
                memcpy(&m, p + i * row + dim * sb, 4);
                massive = m != 0.0f;
            }
            if (massive) idx.push_back((uint32_t)i);
        }
        s->subset = idx.size() != n;
    }
    s->n_affecting = s->subset ? idx.size() : n;
    cudaError_t e = cudaSuccess;
    if (n) {
        e = s->particles.ensure(n * row);
        if (e == cudaSuccess) e = s->vel.ensure(n * dim * sb);
        if (e == cudaSuccess) e = s->acc.ensure(n * dim * sb);
        if (e == cudaSuccess && s->subset && !idx.empty()) {
            e = s->idx.ensure(idx.size() * sizeof(uint32_t));
            if (e == cudaSuccess) e = s->src.ensure(idx.size() * row);
        }
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(s->particles.p, particles, n * row, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            e = velocities ? cudaMemcpyAsync(s->vel.p, velocities, n * dim * sb, cudaMemcpyHostToDevice,
                                             ctx->stream)
                           : cudaMemsetAsync(s->vel.p, 0, n * dim * sb, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(s->acc.p, 0, n * dim * sb, ctx->stream);
        if (e == cudaSuccess && s->subset && !idx.empty())
            e = cudaMemcpyAsync(s->idx.p, idx.data(), idx.size() * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // host buffers are borrowed for the call only
    }
    if (e != cudaSuccess)
        return bail(fail(ctx, e == cudaErrorMemoryAllocation ? PCUDA_ERR_OUT_OF_MEMORY : PCUDA_ERR_CUDA,
                         "pcuda_sim_create: %s", cudaGetErrorString(e)));
    ctx->live_sims++;
    *out = s;
    return PCUDA_OK;
}

int pcuda_sim_configure(pcuda_ctx *ctx, pcuda_sim *s, const pcuda_sim_config *config) {
    if (!ctx || !s || !config) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    PCUDA_TRY(sim::validate(ctx, config));
    if (config->dim != s->cfg.dim || config->scalar != s->cfg.scalar ||
        (config->flags & PCUDA_SIM_AFFECTING_MASSIVE_ONLY) != (s->cfg.flags & PCUDA_SIM_AFFECTING_MASSIVE_ONLY))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "dim, scalar and the affecting subset are fixed at creation");
    DeviceGuard guard(ctx->device);
    s->cfg = *config;
    s->warm = false;  // dt / softening / algorithm are baked into a captured graph
    sim::drop_graph(s);
    return PCUDA_OK;
}

int pcuda_sim_step(pcuda_ctx *ctx, pcuda_sim *s, uint32_t n_steps) {
    if (!ctx || !s) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(ctx->device);
    ctx->launches = 0;
    if (s->n == 0) {
        s->steps_done += n_steps;
        return PCUDA_OK;
    }
    uint32_t left = n_steps;
    bool graphable = s->cfg.algorithm == PCUDA_BRUTE_FORCE && !(s->cfg.flags & PCUDA_SIM_NO_GRAPH);
    while (left) {
        // the first step runs eagerly: it sizes every scratch buffer and counts its launches
        if (graphable && s->warm && left >= (uint32_t)sim::GRAPH_STEPS) {
            const void *key[5];
            sim::scratch_key(ctx, key);
            if ((!s->graph || memcmp(key, s->graph_key, sizeof key) != 0) && !sim::capture(ctx, s)) {
                graphable = false;  // not capturable here: stay eager for good
                s->cfg.flags |= PCUDA_SIM_NO_GRAPH;
                continue;
            }
            PCUDA_CUDA_TRY(ctx, cudaGraphLaunch(s->graph, ctx->stream));
            ctx->launches += s->launches_per_step * sim::GRAPH_STEPS;
            s->steps_done += sim::GRAPH_STEPS;
            left -= sim::GRAPH_STEPS;
            continue;
        }
        const uint32_t before = ctx->launches;
        PCUDA_TRY(sim::step_once(ctx, s));
        s->launches_per_step = ctx->launches - before;
        s->warm = true;
        s->steps_done++;
        --left;
    }
    ctx->timings.kernel_launches = ctx->launches;
    return PCUDA_OK;
}

int pcuda_sim_read(pcuda_ctx *ctx, pcuda_sim *s, void *particles, void *velocities,
                   void *accelerations) {
    if (!ctx || !s) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(ctx->device);
    const size_t sb = s->scalar_bytes(), dim = s->cfg.dim, n = s->n;
    if (n) {
        if (particles)
            PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(particles, s->particles.p, n * (dim + 1) * sb,
                                                cudaMemcpyDeviceToHost, ctx->stream));
        if (velocities)
            PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(velocities, s->vel.p, n * dim * sb,
                                                cudaMemcpyDeviceToHost, ctx->stream));
        if (accelerations)
            PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(accelerations, s->acc.p, n * dim * sb,
                                                cudaMemcpyDeviceToHost, ctx->stream));
    }
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PCUDA_OK;
}

int pcuda_sim_info(const pcuda_sim *s, pcuda_sim_info_t *out) {
    if (!s || !out) return PCUDA_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof *out);
    out->n_particles = s->n;
    out->n_affecting = s->n_affecting;
    out->steps_done = s->steps_done;
    out->d_particles = s->particles.p;
    out->d_velocities = s->vel.p;
    out->d_accelerations = s->acc.p;
    out->graph_active = s->graph != nullptr;
    out->launches_per_step = s->launches_per_step;
    return PCUDA_OK;
}

void pcuda_sim_destroy(pcuda_ctx *ctx, pcuda_sim *s) {
    if (!s) return;
    if (ctx) {
        DeviceGuard guard(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->live_sims > 0) ctx->live_sims--;
    }
    sim::drop_graph(s);
    DevBuf *bufs[] = {&s->particles, &s->vel, &s->acc, &s->idx, &s->src};
    for (DevBuf *b : bufs) b->release();
    delete s;
}

}  // extern "C"
