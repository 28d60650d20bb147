// bruteforce.cu — K1: softened pairwise gravity, O(n_affected * n_affecting), sm_100a.
//
// Replaces the reference's wgpu operator gpu::BruteForce (particular/src/gpu/mod.rs:179-208 and
// the two WGSL templates gpu/bruteforce*.wgsl); the arithmetic follows the scalar pair kernel
// gravity/impls/mod.rs:151-166 evaluated per (affected, affecting) pair as sequential::BruteForce
// does (sequential.rs:181-194).  It is a new design, not a translation of the shaders:
//
//   * sources are streamed as 16-byte {x,y,z,mu} records in tiles that the TMA engine bulk-copies
//     (cp.async.bulk + mbarrier complete_tx) into a 4-stage shared-memory ring; consumers read a
//     record with one broadcast LDS.128;
//   * every thread owns 2*TP targets held in registers as packed pairs, and all pair arithmetic
//     is issued as packed FP32 (FADD2 / FFMA2 / FMUL2): 12 packed instructions + 2 MUFU.RSQ per
//     two pairs, so the FP32 pipe (not the issue slots) is the limiter;
//   * the grid is (target tiles) x (source splits) so that tens of waves of equal-cost CTAs cover
//     the 148 SMs; split partial sums are reduced in a fixed order (deterministic results).
//
// FP32 work: 20 flop / pair by the GPU-Gems-3 convention the reference cites
// (gpu/resources.rs:76-77).  Bytes are irrelevant: the source set lives in L2.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "ptx.cuh"

namespace pcuda {
namespace bf {

constexpr int STAGES = 4;
constexpr int PREFETCH = 2;     // tiles in flight ahead of the consumers
constexpr int TILE_MAX = 256;   // source records per stage
constexpr float TINY_R2 = 1e-18f;  // clamp variant of CHECKED && eps == 0 (tuning only)
constexpr float PAD_POS = 1e18f;   // zero-mass padding records sit here: they contribute exactly 0

// ------------------------------------------------------------------------------------------------
// f32 kernel.  Source record: DIM 3 -> {x,y,z,mu}; DIM 2 -> {x,y,mu,0}.
// CLAMP != 0 <=> (checked && eps*eps == 0): a coincident pair must contribute nothing
// (impls/mod.rs:160-161).
//   CLAMP == 1: r2 == 0 is replaced by +inf, so rsqrt gives 0 and the term is d * 0 = 0 for any
//               finite mu.  Exact; two ALU-pipe instructions per pair (FSETP + FSEL), ~5 % slower.
//   CLAMP == 2: r2 is clamped from below with one FMNMX to t = 2 (max|mu| * 1e-38)^(2/3), derived
//               from the largest |mu| of the call (mass_max_kernel) so that mu * r^-3 stays
//               finite; a coincident pair then gives d * finite = 0 exactly.  (Tuning only.)
//   CLAMP == 3: the same threshold t is ADDED to r2 through the FMA chain that already adds
//               eps^2 (zero extra instructions; every instruction issued next to the FMA pipe
//               costs about one pipe cycle on sm_100).  Coincident pairs give
//               exactly 0; r2 + t == r2 bit for bit whenever r2 >= 2^24 t, i.e. for separations
//               above ~1.3e-6 * (max|mu| / 1e9)^(1/3), and t ~ 1e-19 is 9 orders of magnitude
//               below the f32 resolution of positions of magnitude >= 1e-3.  Default for large
//               problems.
// CLAMP == 0: either eps2 > 0 (d = 0 gives 0 without any test) or the caller asked for the
// unchecked reference behaviour (coincident pair -> NaN, as in the reference).
template <int DIM, int TP, int BLOCK, int MINB, int CLAMP, bool FUSE>
__global__ void __launch_bounds__(BLOCK, MINB)
    pair_kernel_f32(const float *__restrict__ tgt, int tgt_stride, int n_tgt,
                    const float4 *__restrict__ src, int n_src, int src_chunk, int tile, float eps2,
                    float *__restrict__ out, float *__restrict__ partial, size_t n_pad,
                    const unsigned *__restrict__ mass_max_bits, unsigned *__restrict__ tile_done) {
    __shared__ __align__(128) float4 tiles[STAGES][TILE_MAX];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];

    const int tid = threadIdx.x;
    const int lane = tid & 31;

    // ---- targets: 2*TP per thread, negated, packed as (even, odd) pairs ----
    const int tbase = blockIdx.x * (BLOCK * 2 * TP) + tid;
    float2 ntx[TP], nty[TP], ntz[TP];
#pragma unroll
    for (int p = 0; p < TP; ++p) {
        int i0 = min(tbase + (2 * p) * BLOCK, n_tgt - 1);
        int i1 = min(tbase + (2 * p + 1) * BLOCK, n_tgt - 1);
        const float *a = tgt + (size_t)i0 * tgt_stride;
        const float *b = tgt + (size_t)i1 * tgt_stride;
        ntx[p] = make_float2(-a[0], -b[0]);
        nty[p] = make_float2(-a[1], -b[1]);
        if (DIM == 3) ntz[p] = make_float2(-a[2], -b[2]);
    }
    float2 ax[TP], ay[TP], az[TP];
#pragma unroll
    for (int p = 0; p < TP; ++p) ax[p] = ay[p] = az[p] = make_float2(0.f, 0.f);

    const int s_begin = blockIdx.y * src_chunk;
    const int s_end = min(n_src, s_begin + src_chunk);
    const int ntiles = (s_end - s_begin + tile - 1) / tile;

    auto issue = [&](int t) {  // executed by thread 0 only
        const int st = t % STAGES;
        const int first = s_begin + t * tile;
        const int cnt = min(tile, s_end - first);
        const int cnt4 = (cnt + 3) & ~3;
        for (int i = cnt; i < cnt4; ++i)
            tiles[st][i] = DIM == 3 ? make_float4(PAD_POS, PAD_POS, PAD_POS, 0.f)
                                    : make_float4(PAD_POS, PAD_POS, 0.f, 0.f);
        ptx::mbar_arrive_expect_tx(&full_bar[st], (uint32_t)cnt * 16u);
        ptx::bulk_g2s(&tiles[st][0], src + first, (uint32_t)cnt * 16u, &full_bar[st]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], BLOCK / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < PREFETCH && t < ntiles; ++t) issue(t);

    float2 eps2p = make_float2(eps2, eps2);
    float tiny_r2 = TINY_R2;
    if (CLAMP >= 2) {
        // smallest r2 for which max|mu| * r2^-1.5 < ~1e38:  r2 > (max|mu| * 1e-38)^(2/3)
        const float mmax = __uint_as_float(*mass_max_bits);
        const float c = cbrtf(fminf(mmax, 3e38f)) * 2.2e-13f;  // 2.2e-13 ~ cbrt(1e-38)
        tiny_r2 = fmaxf(2.f * c * c, 1e-36f);
        if (CLAMP == 3) eps2p = make_float2(tiny_r2, tiny_r2);
    }

    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + PREFETCH < ntiles) {
            const int tn = t + PREFETCH;
            if (tn >= STAGES) ptx::mbar_wait(&empty_bar[tn % STAGES], ((tn / STAGES) - 1) & 1);
            issue(tn);
        }
        const int st = t % STAGES;
        ptx::mbar_wait(&full_bar[st], (t / STAGES) & 1);

        const int cnt4 = (min(tile, s_end - (s_begin + t * tile)) + 3) & ~3;
        const float4 *sp = tiles[st];
#pragma unroll 1
        for (int j = 0; j < cnt4; j += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 s = sp[j + u];
                const float sm = DIM == 3 ? s.w : s.z;
#pragma unroll
                for (int p = 0; p < TP; ++p) {
                    const float2 dx = ptx::add2(ntx[p], ptx::splat(s.x));
                    const float2 dy = ptx::add2(nty[p], ptx::splat(s.y));
                    float2 r2 = ptx::fma2(dx, dx, eps2p);
                    r2 = ptx::fma2(dy, dy, r2);
                    float2 dz;
                    if (DIM == 3) {
                        dz = ptx::add2(ntz[p], ptx::splat(s.z));
                        r2 = ptx::fma2(dz, dz, r2);
                    }
                    if (CLAMP == 1) {
                        r2.x = r2.x == 0.f ? __int_as_float(0x7f800000) : r2.x;
                        r2.y = r2.y == 0.f ? __int_as_float(0x7f800000) : r2.y;
                    } else if (CLAMP == 2) {
                        r2.x = fmaxf(r2.x, tiny_r2);
                        r2.y = fmaxf(r2.y, tiny_r2);
                    }
                    float2 ri;
                    ri.x = ptx::rsqrt_approx(r2.x);
                    ri.y = ptx::rsqrt_approx(r2.y);
                    const float2 ri2 = ptx::mul2(ri, ri);
                    const float2 mri = ptx::mul2(ri, ptx::splat(sm));
                    const float2 sc = ptx::mul2(ri2, mri);
                    ax[p] = ptx::fma2(dx, sc, ax[p]);
                    ay[p] = ptx::fma2(dy, sc, ay[p]);
                    if (DIM == 3) az[p] = ptx::fma2(dz, sc, az[p]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty_bar[st]);
    }

    // ---- results ----
    // One source split: straight to `out`.  Several: every CTA stores its partial sums.  FUSE (small
    // problems): the CTA that finishes LAST for a target tile (a ticket per tile) adds the partials
    // of all splits in split order — a fixed order, so the result does not depend on which CTA that
    // was — and writes `out`; no second kernel, because at the reference's criterion sizes a
    // dependent launch costs as much as the whole evaluation.  Large problems keep the separate
    // reduce_partials kernel (same order, same bits): the ticket epilogue costs the hot loop's
    // register allocation 0.7 % there.
    const bool direct = gridDim.y == 1;
#pragma unroll
    for (int p = 0; p < TP; ++p) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = tbase + (2 * p + h) * BLOCK;
            if (i >= n_tgt) continue;
            const float vx = h ? ax[p].y : ax[p].x;
            const float vy = h ? ay[p].y : ay[p].x;
            const float vz = DIM == 3 ? (h ? az[p].y : az[p].x) : 0.f;
            if (direct) {
                out[(size_t)i * DIM + 0] = vx;
                out[(size_t)i * DIM + 1] = vy;
                if (DIM == 3) out[(size_t)i * DIM + 2] = vz;
            } else {
                float *pp = partial + (size_t)blockIdx.y * DIM * n_pad + i;
                pp[0] = vx;
                pp[n_pad] = vy;
                if (DIM == 3) pp[2 * n_pad] = vz;
            }
        }
    }
    if (direct || !FUSE) return;  // !FUSE: reduce_partials runs as a second kernel
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&tile_done[blockIdx.x], 1u) == gridDim.y - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int splits = (int)gridDim.y;
    for (int k = 0; k < 2 * TP; ++k) {
        const int i = tbase + k * BLOCK;
        if (i >= n_tgt) break;
        // eight splits x DIM components of loads in flight at a time (they are L2 round trips);
        // the additions stay in split order
        float acc[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc[c] = 0.f;
        for (int y0 = 0; y0 < splits; y0 += 8) {
            float v[8][DIM];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                    v[u][c] = y0 + u < splits ? __ldcg(partial + ((size_t)(y0 + u) * DIM + c) * n_pad + i) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                    if (y0 + u < splits) acc[c] += v[u][c];
        }
#pragma unroll
        for (int c = 0; c < DIM; ++c) out[(size_t)i * DIM + c] = acc[c];
    }
    if (tid == 0) tile_done[blockIdx.x] = 0;  // ready for the next launch
}

// max |mu| over the source records (bit pattern of a non-negative float orders like an unsigned).
__global__ void __launch_bounds__(256) mass_max_kernel(const float4 *__restrict__ src, int n, int dim,
                                                       unsigned *__restrict__ out) {
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 s = src[i];
        m = fmaxf(m, fabsf(dim == 3 ? s.w : s.z));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// Fixed-order reduction of the source-split partial sums: out[i][c] = sum_y partial[y][c][i].
template <typename T, int DIM>
__global__ void reduce_partials(const T *__restrict__ partial, int splits, size_t n_pad, int n_tgt,
                                T *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tgt) return;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        T acc = 0;
        for (int y = 0; y < splits; ++y) acc += partial[((size_t)y * DIM + c) * n_pad + i];
        out[(size_t)i * DIM + c] = acc;
    }
}

// DIM 2 sources arrive as {x,y,mu} (12 B); the kernel wants 16-byte {x,y,mu,0} records.
__global__ void pack_sources_2d(const float *__restrict__ in, int n, float4 *__restrict__ outp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) outp[i] = make_float4(in[3 * (size_t)i], in[3 * (size_t)i + 1], in[3 * (size_t)i + 2], 0.f);
}

// ------------------------------------------------------------------------------------------------
// f64 kernel (the "precision path", BASELINE config 5a).  Scalar DFMA arithmetic, T targets per
// thread, same TMA ring with 32-byte {x,y,z,mu} records.  rsqrt(double) is CUDA's <= 1 ulp
// implementation (MUFU.RSQ64H + Newton steps); per-term error is a few 1e-16, far inside the
// 1e-12 parity bound.
constexpr int TILE64 = 128;
constexpr double PAD_POS_64 = 1e100;

using ptx::mu_rcbrt2;
using ptx::one_if_zero;

// DIM 2 (DVec2): sources repacked to {x, y, 0, mu}; the z terms are compiled out.
template <int DIM, int T, int BLOCK, bool CLAMP>
__global__ void __launch_bounds__(BLOCK)
    pair_kernel_f64(const double *__restrict__ tgt, int tgt_stride, int n_tgt,
                    const double4 *__restrict__ src, int n_src, int src_chunk, int tile,
                    double eps2, double *__restrict__ out, double *__restrict__ partial,
                    size_t n_pad) {
    __shared__ __align__(128) double4 tiles[STAGES][TILE64];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tbase = blockIdx.x * (BLOCK * T) + tid;
    double tx[T], ty[T], tz[T], ax[T], ay[T], az[T];
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const int i = min(tbase + k * BLOCK, n_tgt - 1);
        const double *a = tgt + (size_t)i * tgt_stride;
        tx[k] = a[0];
        ty[k] = a[1];
        tz[k] = DIM == 3 ? a[2] : 0.0;
        ax[k] = ay[k] = az[k] = 0.0;
    }

    const int s_begin = blockIdx.y * src_chunk;
    const int s_end = min(n_src, s_begin + src_chunk);
    const int ntiles = (s_end - s_begin + tile - 1) / tile;

    auto issue = [&](int t) {
        const int st = t % STAGES;
        const int first = s_begin + t * tile;
        const int cnt = min(tile, s_end - first);
        const int cnt2 = (cnt + 1) & ~1;
        for (int i = cnt; i < cnt2; ++i)
            tiles[st][i] = make_double4(PAD_POS_64, PAD_POS_64, PAD_POS_64, 0.0);
        ptx::mbar_arrive_expect_tx(&full_bar[st], (uint32_t)cnt * 32u);
        ptx::bulk_g2s(&tiles[st][0], src + first, (uint32_t)cnt * 32u, &full_bar[st]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], BLOCK / 32);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < PREFETCH && t < ntiles; ++t) issue(t);

    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + PREFETCH < ntiles) {
            const int tn = t + PREFETCH;
            if (tn >= STAGES) ptx::mbar_wait(&empty_bar[tn % STAGES], ((tn / STAGES) - 1) & 1);
            issue(tn);
        }
        const int st = t % STAGES;
        ptx::mbar_wait(&full_bar[st], (t / STAGES) & 1);
        const int cnt2 = (min(tile, s_end - (s_begin + t * tile)) + 1) & ~1;
        const double4 *sp = tiles[st];
#pragma unroll 1
        for (int j = 0; j < cnt2; j += 2) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double4 s = sp[j + u];
#pragma unroll
                for (int k = 0; k < T; ++k) {
                    const double dx = s.x - tx[k], dy = s.y - ty[k];
                    double r2 = fma(dx, dx, eps2);
                    r2 = fma(dy, dy, r2);
                    double dz = 0.0;
                    if (DIM == 3) {
                        dz = s.z - tz[k];
                        r2 = fma(dz, dz, r2);
                    }
                    if (CLAMP) r2 = one_if_zero(r2);  // d == 0 there, so the term is 0 * finite = 0
                    const double sc = mu_rcbrt2(r2, s.w);
                    ax[k] = fma(dx, sc, ax[k]);
                    ay[k] = fma(dy, sc, ay[k]);
                    if (DIM == 3) az[k] = fma(dz, sc, az[k]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty_bar[st]);
    }

    const bool direct = gridDim.y == 1;
#pragma unroll
    for (int k = 0; k < T; ++k) {
        const int i = tbase + k * BLOCK;
        if (i >= n_tgt) continue;
        if (direct) {
            out[(size_t)i * DIM + 0] = ax[k];
            out[(size_t)i * DIM + 1] = ay[k];
            if (DIM == 3) out[(size_t)i * DIM + 2] = az[k];
        } else {
            double *pp = partial + (size_t)blockIdx.y * DIM * n_pad + i;
            pp[0] = ax[k];
            pp[n_pad] = ay[k];
            if (DIM == 3) pp[2 * n_pad] = az[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Launch planning.  A CTA covers `tile_t` targets and one chunk of sources; we want enough
// equal-cost CTAs for >= ~16 waves over SMs x resident CTAs (tail < 6 %), but chunks of at least
// a few tiles so the pipeline prologue is amortised.
static int g_force_tp = 0;    // test / tuning hooks (pcuda_debug_set)
static int g_clamp_mode = 0;   // 0 = automatic (by problem size), 1 = select, 2 = clamp, 3 = additive
static int g_waves = 48;       // equal-cost CTAs per resident slot (tail < 1 / g_waves)

struct Plan {
    int tp, block, minb;
    int n_tb, splits, chunk, tile;
    bool small;  // latency-bound problem size (see make_plan)
};

static Plan make_plan(int sm_count, size_t na, size_t nb, int force_tp) {
    Plan pl{};
    // Targets per thread.  Large problems (many sources): 8 targets per thread as soon as there are a
    // few dozen target tiles — the source splits supply the CTAs — because one LDS.128 then feeds four
    // packed pairs (B200, 1M sources: 125k targets 0.690 of peak against 0.688 with 2 per thread and a
    // 1.6 % loss to the last, nearly empty tile that run_f32 now gives to a second launch; 16k targets:
    // 2 per thread wins, 0.675 against 0.669).  Small problems keep the occupancy-driven choice.
    const bool large = (double)na * (double)nb >= 2.5e8;
    if (force_tp > 0) pl.tp = force_tp;
    else if (large && na >= 49152) pl.tp = 4;
    else if (!large && na >= (size_t)sm_count * 2 * 2048) pl.tp = 4;
    else if (!large && na >= (size_t)sm_count * 3 * 512) pl.tp = 2;
    else pl.tp = 1;
    pl.block = pl.tp == 1 ? 128 : 256;
    pl.minb = pl.tp == 4 ? 2 : (pl.tp == 2 ? 3 : 4);
    const int tile_t = pl.block * 2 * pl.tp;
    pl.n_tb = (int)((na + tile_t - 1) / tile_t);
    // Small problems (the reference's criterion sizes) are latency bound: a warp retires one source
    // per ~70 cycles, so the sources are cut as finely as one 64-record tile per CTA and two waves
    // of CTAs are enough; large problems want tens of waves of long, equal-cost CTAs.
    const bool small = (double)na * (double)nb < 2.5e8;
    pl.small = small;
    pl.tile = !small && nb >= 4096 ? TILE_MAX : 64;
    const long slots = (long)sm_count * pl.minb;
    const long want = slots * (small ? 2 : g_waves);
    long splits = (want + pl.n_tb - 1) / pl.n_tb;
    const long max_splits = std::max<long>(1, (long)(nb / ((size_t)pl.tile * (small ? 1 : 4))));
    splits = std::max<long>(1, std::min(splits, max_splits));
    if (splits > 1 && pl.n_tb >= want) splits = 1;
    long chunk = ((long)nb + splits - 1) / splits;
    chunk = ((chunk + pl.tile - 1) / pl.tile) * pl.tile;
    splits = ((long)nb + chunk - 1) / chunk;
    pl.splits = (int)std::max<long>(1, splits);
    pl.chunk = (int)chunk;
    return pl;
}


template <int DIM, int TP, int BLOCK, int MINB, bool FUSE>
static cudaError_t launch_f32(const Plan &pl, int clamp, cudaStream_t stream, const float *tgt,
                              int tgt_stride, int na, const float4 *src, int nb, float eps2,
                              float *out, float *partial, size_t n_pad, const unsigned *mass_max,
                              unsigned *tile_done) {
    dim3 grid(pl.n_tb, pl.splits);
    if (clamp == 3)
        pair_kernel_f32<DIM, TP, BLOCK, MINB, 3, FUSE><<<grid, BLOCK, 0, stream>>>(
            tgt, tgt_stride, na, src, nb, pl.chunk, pl.tile, eps2, out, partial, n_pad, mass_max, tile_done);
    else if (clamp == 2)
        pair_kernel_f32<DIM, TP, BLOCK, MINB, 2, FUSE><<<grid, BLOCK, 0, stream>>>(
            tgt, tgt_stride, na, src, nb, pl.chunk, pl.tile, eps2, out, partial, n_pad, mass_max, tile_done);
    else if (clamp == 1)
        pair_kernel_f32<DIM, TP, BLOCK, MINB, 1, FUSE><<<grid, BLOCK, 0, stream>>>(
            tgt, tgt_stride, na, src, nb, pl.chunk, pl.tile, eps2, out, partial, n_pad, mass_max, tile_done);
    else
        pair_kernel_f32<DIM, TP, BLOCK, MINB, 0, FUSE><<<grid, BLOCK, 0, stream>>>(
            tgt, tgt_stride, na, src, nb, pl.chunk, pl.tile, eps2, out, partial, n_pad, mass_max, tile_done);
    return cudaGetLastError();
}

// Per-target-tile tickets of the "last CTA reduces" epilogue: zero when allocated, and every
// launch leaves them zero again.
static int ensure_tile_tickets(pcuda_ctx *ctx, size_t n_tiles) {
    const size_t bytes = n_tiles * sizeof(unsigned);
    if (bytes <= ctx->d_tile_done.cap) return PCUDA_OK;
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // a launch in flight may still hold tickets
    PCUDA_CUDA_TRY(ctx, ctx->d_tile_done.ensure(bytes));
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_tile_done.p, 0, ctx->d_tile_done.cap, ctx->stream));
    return PCUDA_OK;
}

// One launch (+ reduction) of the pair kernel over `na` targets with the plan of (na, nb).
template <int DIM>
static int run_f32_part(pcuda_ctx *ctx, const float *d_tgt, int tgt_stride, size_t na, const float4 *d_src4,
                        size_t nb, float eps2, int clamp, const unsigned *mass_max, float *d_out,
                        int force_tp) {
    const Plan pl = make_plan(ctx->sm_count, na, nb, force_tp);
    const size_t n_pad = (na + 63) & ~size_t(63);
    float *partial = nullptr;
    unsigned *tile_done = nullptr;
    const bool fuse = pl.small;  // the last CTA of a target tile reduces the split partial sums
    if (pl.splits > 1) {
        PCUDA_CUDA_TRY(ctx, ctx->d_partial.ensure((size_t)pl.splits * DIM * n_pad * sizeof(float)));
        partial = ctx->d_partial.as<float>();
        if (fuse) {
            PCUDA_TRY(ensure_tile_tickets(ctx, (size_t)pl.n_tb));
            tile_done = ctx->d_tile_done.as<unsigned>();
        }
    }
    cudaError_t e;
    if (pl.tp == 4)
        e = launch_f32<DIM, 4, 256, 2, false>(pl, clamp, ctx->stream, d_tgt, tgt_stride, (int)na, d_src4,
                                              (int)nb, eps2, d_out, partial, n_pad, mass_max, tile_done);
    else if (pl.tp == 2)
        e = fuse ? launch_f32<DIM, 2, 256, 3, true>(pl, clamp, ctx->stream, d_tgt, tgt_stride, (int)na, d_src4,
                                                    (int)nb, eps2, d_out, partial, n_pad, mass_max, tile_done)
                 : launch_f32<DIM, 2, 256, 3, false>(pl, clamp, ctx->stream, d_tgt, tgt_stride, (int)na, d_src4,
                                                     (int)nb, eps2, d_out, partial, n_pad, mass_max, tile_done);
    else
        e = fuse ? launch_f32<DIM, 1, 128, 4, true>(pl, clamp, ctx->stream, d_tgt, tgt_stride, (int)na, d_src4,
                                                    (int)nb, eps2, d_out, partial, n_pad, mass_max, tile_done)
                 : launch_f32<DIM, 1, 128, 4, false>(pl, clamp, ctx->stream, d_tgt, tgt_stride, (int)na, d_src4,
                                                     (int)nb, eps2, d_out, partial, n_pad, mass_max, tile_done);
    PCUDA_CUDA_TRY(ctx, e);
    ctx->launches++;
    if (pl.splits > 1 && !(fuse && pl.tp != 4)) {
        reduce_partials<float, DIM><<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(
            partial, pl.splits, n_pad, (int)na, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return PCUDA_OK;
}

// Enqueues the brute-force evaluation on ctx->stream.  All pointers are device pointers.
// src4: 16-byte records (DIM 3: {x,y,z,mu}; DIM 2: {x,y,mu,0}); tgt rows have tgt_stride floats.
template <int DIM>
static int run_f32(pcuda_ctx *ctx, const float *d_tgt, int tgt_stride, size_t na,
                   const float4 *d_src4, size_t nb, float softening, int checked, float *d_out) {
    if (na == 0) return PCUDA_OK;
    if (nb == 0) {
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, na * DIM * sizeof(float), ctx->stream));
        return PCUDA_OK;
    }
    if (na > 0x7fffffffull || nb > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    const float eps2 = softening * softening;
    int clamp = 0;
    if (checked && eps2 == 0.0f)
        clamp = g_clamp_mode ? g_clamp_mode
                             : ((double)na * (double)nb >= 2.5e8 && !ctx->exact_checked ? 3 : 1);
    unsigned *mass_max = nullptr;
    if (clamp >= 2) {
        PCUDA_CUDA_TRY(ctx, ctx->d_massmax.ensure(sizeof(unsigned)));
        mass_max = ctx->d_massmax.as<unsigned>();
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(mass_max, 0, sizeof(unsigned), ctx->stream));
        const int blocks = (int)std::min<size_t>((size_t)ctx->sm_count * 4, (nb + 255) / 256);
        mass_max_kernel<<<blocks, 256, 0, ctx->stream>>>(d_src4, (int)nb, DIM, mass_max);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    // The 8-targets-per-thread variant works in tiles of 2048 targets.  A last tile that is mostly
    // empty costs a full tile's time (125 000 targets = 61.04 tiles: 1.6 % of the launch), so the
    // remainder goes to a second launch with its own plan (finer tiles, its own source splits).
    const Plan pl = make_plan(ctx->sm_count, na, nb, g_force_tp);
    const size_t tile_t = (size_t)pl.block * 2 * pl.tp;
    const size_t rem = na % tile_t;
    if (pl.tp == 4 && !pl.small && g_force_tp == 0 && na > tile_t && rem != 0 && rem <= tile_t / 2) {
        const size_t head = na - rem;
        PCUDA_TRY(run_f32_part<DIM>(ctx, d_tgt, tgt_stride, head, d_src4, nb, eps2, clamp, mass_max, d_out, 4));
        return run_f32_part<DIM>(ctx, d_tgt + head * tgt_stride, tgt_stride, rem, d_src4, nb, eps2, clamp,
                                 mass_max, d_out + head * DIM, 0);
    }
    return run_f32_part<DIM>(ctx, d_tgt, tgt_stride, na, d_src4, nb, eps2, clamp, mass_max, d_out, g_force_tp);
}

template <int DIM = 3>
static int run_f64(pcuda_ctx *ctx, const double *d_tgt, int tgt_stride, size_t na,
                   const double4 *d_src4, size_t nb, double softening, int checked,
                   double *d_out) {
    if (na == 0) return PCUDA_OK;
    if (nb == 0) {
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, na * DIM * sizeof(double), ctx->stream));
        return PCUDA_OK;
    }
    if (na > 0x7fffffffull || nb > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    const double eps2 = softening * softening;
    const bool clamp = checked && eps2 == 0.0;
    constexpr int T = 4, BLOCK = 128;
    const int tile_t = T * BLOCK;
    const int n_tb = (int)((na + tile_t - 1) / tile_t);
    const int tile = nb >= 2048 ? TILE64 : 32;
    const long want = (long)ctx->sm_count * 4 * 16;
    long splits = std::max<long>(1, std::min<long>((want + n_tb - 1) / n_tb,
                                                   std::max<long>(1, (long)(nb / ((size_t)tile * 4)))));
    if (n_tb >= want) splits = 1;
    long chunk = ((long)nb + splits - 1) / splits;
    chunk = ((chunk + tile - 1) / tile) * tile;
    splits = ((long)nb + chunk - 1) / chunk;
    const size_t n_pad = (na + 63) & ~size_t(63);
    double *partial = nullptr;
    if (splits > 1) {
        PCUDA_CUDA_TRY(ctx, ctx->d_partial.ensure((size_t)splits * DIM * n_pad * sizeof(double)));
        partial = ctx->d_partial.as<double>();
    }
    dim3 grid(n_tb, (unsigned)splits);
    if (clamp)
        pair_kernel_f64<DIM, T, BLOCK, true><<<grid, BLOCK, 0, ctx->stream>>>(
            d_tgt, tgt_stride, (int)na, d_src4, (int)nb, (int)chunk, tile, eps2, d_out, partial, n_pad);
    else
        pair_kernel_f64<DIM, T, BLOCK, false><<<grid, BLOCK, 0, ctx->stream>>>(
            d_tgt, tgt_stride, (int)na, d_src4, (int)nb, (int)chunk, tile, eps2, d_out, partial, n_pad);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (splits > 1) {
        reduce_partials<double, DIM><<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(
            partial, (int)splits, n_pad, (int)na, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return PCUDA_OK;
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Device-pointer entry: DIM 3.
static int dev_f32x3(pcuda_ctx *ctx, const float *d_aff, size_t na, const float *d_src, size_t nb,
                     float eps, int checked, float *d_out) {
    if (nb && !aligned(d_src, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "affecting must be 16-byte aligned");
    const float *tgt = d_aff ? d_aff : d_src;
    const int stride = d_aff ? 3 : 4;
    return run_f32<3>(ctx, tgt, stride, na, reinterpret_cast<const float4 *>(d_src), nb, eps,
                      checked, d_out);
}

// tgt == nullptr: the targets are the sources (stride 3).
static int run_f32x2(pcuda_ctx *ctx, const float *d_tgt, int tgt_stride, size_t na,
                     const float *d_src, size_t nb, float eps, int checked, float *d_out) {
    float4 *packed = nullptr;
    if (nb) {
        PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure(nb * sizeof(float4)));
        packed = ctx->d_packed_src.as<float4>();
        pack_sources_2d<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(d_src, (int)nb, packed);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return run_f32<2>(ctx, d_tgt ? d_tgt : d_src, d_tgt ? tgt_stride : 3, na, packed, nb, eps,
                      checked, d_out);
}

static int dev_f32x2(pcuda_ctx *ctx, const float *d_aff, size_t na, const float *d_src, size_t nb,
                     float eps, int checked, float *d_out) {
    return run_f32x2(ctx, d_aff, 2, na, d_src, nb, eps, checked, d_out);
}

// DVec2: {x,y,mu} (24 B) -> {x, y, 0, mu} records for the f64 kernel.
__global__ void pack_sources_2d_f64(const double *__restrict__ in, int n, double4 *__restrict__ outp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) outp[i] = make_double4(in[3 * (size_t)i], in[3 * (size_t)i + 1], 0.0, in[3 * (size_t)i + 2]);
}

// tgt == nullptr: the targets are the sources (stride 3).
static int run_f64x2(pcuda_ctx *ctx, const double *d_tgt, int tgt_stride, size_t na,
                     const double *d_src, size_t nb, double eps, int checked, double *d_out) {
    double4 *packed = nullptr;
    if (nb) {
        if (nb > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
        PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure(nb * sizeof(double4)));
        packed = ctx->d_packed_src.as<double4>();
        pack_sources_2d_f64<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(d_src, (int)nb, packed);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return run_f64<2>(ctx, d_tgt ? d_tgt : d_src, d_tgt ? tgt_stride : 3, na, packed, nb, eps, checked,
                      d_out);
}

static int dev_f64x2(pcuda_ctx *ctx, const double *d_aff, size_t na, const double *d_src, size_t nb,
                     double eps, int checked, double *d_out) {
    return run_f64x2(ctx, d_aff, 2, na, d_src, nb, eps, checked, d_out);
}

static int dev_f64x3(pcuda_ctx *ctx, const double *d_aff, size_t na, const double *d_src, size_t nb,
                     double eps, int checked, double *d_out) {
    if (nb && !aligned(d_src, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "affecting must be 16-byte aligned");
    const double *tgt = d_aff ? d_aff : d_src;
    const int stride = d_aff ? 3 : 4;
    return run_f64(ctx, tgt, stride, na, reinterpret_cast<const double4 *>(d_src), nb, eps, checked,
                   d_out);
}

// Multi-GPU step (one process per GPU): every rank owns `n_local` particles.  The local records
// are copied into this rank's slot of `d_gathered` (capacity `cap` records per rank; unused slots
// are filled with zero-mass records at PAD_POS, which contribute exactly 0), the slots are
// all-gathered in place over NVLink, and the local targets (aliasing the gathered records,
// stride 4) are evaluated against all world * cap sources.
__global__ void fill_slot_f32x3(const float4 *__restrict__ local, int n_local, int cap,
                                float4 *__restrict__ slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) slot[i] = i < n_local ? local[i] : make_float4(PAD_POS, PAD_POS, PAD_POS, 0.f);
}

static int sharded_f32x3(pcuda_ctx *ctx, const float *d_local, size_t n_local, size_t cap,
                         float eps, int checked, float *d_gathered, float *d_out) {
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    if (cap == 0 || n_local > cap)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "n_local (%zu) exceeds shard capacity (%zu)",
                    n_local, cap);
    if (!aligned(d_local, 16) || !aligned(d_gathered, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "records must be 16-byte aligned");
    if ((size_t)world * cap > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    float4 *all = reinterpret_cast<float4 *>(d_gathered);
    float4 *slot = all + (size_t)rank * cap;
    phase_begin(ctx, PH_COMM);
    fill_slot_f32x3<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const float4 *>(d_local), (int)n_local, (int)cap, slot);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (world > 1) PCUDA_TRY(pcuda_comm_allgather_dev(ctx, slot, all, cap * sizeof(float4)));
    phase_end(ctx, PH_COMM);
    phase_begin(ctx, PH_COMPUTE);
    PCUDA_TRY(run_f32<3>(ctx, reinterpret_cast<const float *>(slot), 4, n_local, all,
                         (size_t)world * cap, eps, checked, d_out));
    phase_end(ctx, PH_COMPUTE);
    return PCUDA_OK;
}

// Multi-GPU `Between(affected, affecting)` step (BASELINE configs[2]: massive -> massless): the
// AFFECTING records are sharded (every rank holds n_local_src of them, capacity `cap` per rank)
// and all-gathered in place into d_gathered; the AFFECTED positions of this rank (its shard of the
// targets, d_aff == nullptr: none) are evaluated against all world * cap sources.  No collective
// touches the targets: with 10k sources the exchange is 160 KB per step.
static int gather_sources_f32x3(pcuda_ctx *ctx, const float *d_local_src, size_t n_local_src,
                                size_t cap, float *d_gathered, size_t *nb_out) {
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    if (cap == 0 || n_local_src > cap)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "n_local (%zu) exceeds shard capacity (%zu)",
                    n_local_src, cap);
    if ((n_local_src && !aligned(d_local_src, 16)) || !aligned(d_gathered, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "records must be 16-byte aligned");
    if ((size_t)world * cap > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    float4 *all = reinterpret_cast<float4 *>(d_gathered);
    float4 *slot = all + (size_t)rank * cap;
    phase_begin(ctx, PH_COMM);
    fill_slot_f32x3<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const float4 *>(d_local_src), (int)n_local_src, (int)cap, slot);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (world > 1) PCUDA_TRY(pcuda_comm_allgather_dev(ctx, slot, all, cap * sizeof(float4)));
    phase_end(ctx, PH_COMM);
    *nb_out = (size_t)world * cap;
    return PCUDA_OK;
}

// Host-pointer wrapper shared by the three precisions: upload, run, download, collect timings.
//
// Large target sets (the massive -> massless split of BASELINE configs[2]: 16M targets, 192 MB up
// and 192 MB down per call) take the chunked path: the targets are cut into chunks whose upload
// (H2D copy engine, own stream), evaluation (context stream) and download (D2H copy engine, own
// stream) overlap through events, with two chunk buffers in flight, so that PCIe disappears behind
// the kernels except for the first upload and the last download.  The reference's wgpu operator
// does the opposite: one blocking write, one dispatch, one blocking map (gpu/resources.rs:318-349).
constexpr size_t CHUNK_MIN_TARGETS = 4u << 20;  // below this the copies are too short to matter
constexpr size_t CHUNK_TARGETS = 1u << 20;

static int ensure_copy_streams(pcuda_ctx *ctx) {
    if (ctx->stream_h2d) return PCUDA_OK;
    PCUDA_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking));
    PCUDA_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk_up[i], cudaEventDisableTiming));
        PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk_done[i], cudaEventDisableTiming));
        PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_chunk_free[i], cudaEventDisableTiming));
    }
    PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_d2h_end, cudaEventDisableTiming));
    return PCUDA_OK;
}

// run(d_tgt, n_tgt, d_src, d_out) enqueues the evaluation of n_tgt targets on ctx->stream.
template <typename S, typename RunFn>
static int host_call_chunked(pcuda_ctx *ctx, const S *affected, size_t na, int dim, S *d_src, S *out,
                             RunFn run) {
    PCUDA_TRY(ensure_copy_streams(ctx));
    const size_t row = (size_t)dim * sizeof(S);
    const size_t chunk = CHUNK_TARGETS;
    // two chunk slots for targets and results
    PCUDA_CUDA_TRY(ctx, ctx->d_affected.ensure(2 * chunk * row));
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(2 * chunk * row));
    S *d_tgt = ctx->d_affected.as<S>(), *d_out = ctx->d_out.as<S>();
    // the copy streams start after everything already queued on the context stream (the source
    // upload in particular)
    PCUDA_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_d2h_end, ctx->stream));
    PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ctx->ev_d2h_end, 0));
    PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_d2h, ctx->ev_d2h_end, 0));
    phase_begin(ctx, PH_COMPUTE);
    size_t k = 0;
    for (size_t off = 0; off < na; off += chunk, ++k) {
        const size_t n = std::min(chunk, na - off);
        const int slot = (int)(k & 1);
        S *t = d_tgt + (size_t)slot * chunk * dim, *o = d_out + (size_t)slot * chunk * dim;
        // slot reuse: the upload of chunk k waits until chunk k-2 has been evaluated, and the
        // evaluation of chunk k until the download of chunk k-2 has left its result slot
        if (k >= 2) PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_h2d, ctx->ev_chunk_done[slot], 0));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(t, affected + off * dim, n * row, cudaMemcpyHostToDevice,
                                            ctx->stream_h2d));
        PCUDA_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_chunk_up[slot], ctx->stream_h2d));
        PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk_up[slot], 0));
        if (k >= 2) PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk_free[slot], 0));
        PCUDA_TRY(run(t, n, d_src, o));
        PCUDA_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_chunk_done[slot], ctx->stream));
        PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream_d2h, ctx->ev_chunk_done[slot], 0));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out + off * dim, o, n * row, cudaMemcpyDeviceToHost,
                                            ctx->stream_d2h));
        PCUDA_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_chunk_free[slot], ctx->stream_d2h));
    }
    phase_end(ctx, PH_COMPUTE);
    // the call ends when the last download has landed
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_d2h_end, ctx->stream_d2h));
    PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h_end, 0));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

template <typename S, typename RunFn>
static int host_targets(pcuda_ctx *ctx, const S *affected, size_t na, int dim, S *d_src, size_t nb,
                        S *out, RunFn run);

template <typename S, typename RunFn>
static int host_call(pcuda_ctx *ctx, const S *affected, size_t na, int dim, const S *affecting,
                     size_t nb, S *out, RunFn run) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if ((na && !out) || (nb && !affecting))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (!affected && na != nb)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "affected == NULL means affected == affecting, but n_affected != n_affecting");
    if (na > 0x7fffffffull || nb > 0x7fffffffull)  // before any buffer is touched
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (na == 0) return PCUDA_OK;
    const size_t src_bytes = nb * (dim + 1) * sizeof(S);
    phase_begin(ctx, PH_UPLOAD);
    S *d_src = nullptr;
    if (nb) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(src_bytes));
        d_src = ctx->d_affecting.as<S>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_src, affecting, src_bytes, cudaMemcpyHostToDevice,
                                            ctx->stream));
    }
    return host_targets<S>(ctx, affected, na, dim, d_src, nb, out, run);
}

// Second half of a host call: the sources are already on the device (or on their way on the
// context stream) and the upload phase is open.  Uploads the targets, evaluates, downloads.
template <typename S, typename RunFn>
static int host_targets(pcuda_ctx *ctx, const S *affected, size_t na, int dim, S *d_src, size_t nb,
                        S *out, RunFn run) {
    const size_t tgt_bytes = na * dim * sizeof(S);
    S *d_tgt = nullptr;
    if (affected && na >= CHUNK_MIN_TARGETS && nb) {
        phase_end(ctx, PH_UPLOAD);
        return host_call_chunked<S>(ctx, affected, na, dim, d_src, out, run);
    }
    if (affected) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affected.ensure(tgt_bytes));
        d_tgt = ctx->d_affected.as<S>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_tgt, affected, tgt_bytes, cudaMemcpyHostToDevice,
                                            ctx->stream));
    }
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(tgt_bytes));
    phase_end(ctx, PH_UPLOAD);
    phase_begin(ctx, PH_COMPUTE);
    PCUDA_TRY(run(d_tgt, na, d_src, ctx->d_out.as<S>()));
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out.p, tgt_bytes, cudaMemcpyDeviceToHost,
                                        ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

template <typename RunFn>
static int dev_call(pcuda_ctx *ctx, RunFn run) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    ctx->launches = 0;
    int s = run();
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

}  // namespace bf

int bf_enqueue_f32(pcuda_ctx *ctx, int dim, const float *d_tgt, int tgt_stride, size_t na,
                   const float *d_src, size_t nb, float softening, int checked, float *d_out) {
    if (dim == 2) return bf::run_f32x2(ctx, d_tgt, tgt_stride, na, d_src, nb, softening, checked, d_out);
    if (nb && !bf::aligned(d_src, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "affecting must be 16-byte aligned");
    return bf::run_f32<3>(ctx, d_tgt, tgt_stride, na, reinterpret_cast<const float4 *>(d_src), nb,
                          softening, checked, d_out);
}

int bf_enqueue_f64(pcuda_ctx *ctx, int dim, const double *d_tgt, int tgt_stride, size_t na,
                   const double *d_src, size_t nb, double softening, int checked, double *d_out) {
    if (dim == 2) return bf::run_f64x2(ctx, d_tgt, tgt_stride, na, d_src, nb, softening, checked, d_out);
    if (nb && !bf::aligned(d_src, 16))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "affecting must be 16-byte aligned");
    return bf::run_f64(ctx, d_tgt, tgt_stride, na, reinterpret_cast<const double4 *>(d_src), nb,
                       softening, checked, d_out);
}

}  // namespace pcuda

using namespace pcuda;

extern "C" {

int pcuda_bruteforce_f32x3(pcuda_ctx *ctx, const float *affected, size_t na, const float *affecting,
                           size_t nb, float softening, int checked, float *out) {
    return bf::host_call<float>(ctx, affected, na, 3, affecting, nb, out,
                                [&](float *dt, size_t n, float *ds, float *dout) {
                                    return bf::dev_f32x3(ctx, dt, n, ds, nb, softening, checked, dout);
                                });
}

int pcuda_bruteforce_f32x2(pcuda_ctx *ctx, const float *affected, size_t na, const float *affecting,
                           size_t nb, float softening, int checked, float *out) {
    return bf::host_call<float>(ctx, affected, na, 2, affecting, nb, out,
                                [&](float *dt, size_t n, float *ds, float *dout) {
                                    return bf::dev_f32x2(ctx, dt, n, ds, nb, softening, checked, dout);
                                });
}

int pcuda_bruteforce_f64x3(pcuda_ctx *ctx, const double *affected, size_t na,
                           const double *affecting, size_t nb, double softening, int checked,
                           double *out) {
    return bf::host_call<double>(ctx, affected, na, 3, affecting, nb, out,
                                 [&](double *dt, size_t n, double *ds, double *dout) {
                                     return bf::dev_f64x3(ctx, dt, n, ds, nb, softening, checked, dout);
                                 });
}

int pcuda_bruteforce_f64x2(pcuda_ctx *ctx, const double *affected, size_t na,
                           const double *affecting, size_t nb, double softening, int checked,
                           double *out) {
    return bf::host_call<double>(ctx, affected, na, 2, affecting, nb, out,
                                 [&](double *dt, size_t n, double *ds, double *dout) {
                                     return bf::dev_f64x2(ctx, dt, n, ds, nb, softening, checked, dout);
                                 });
}

int pcuda_bruteforce_f64x2_dev(pcuda_ctx *ctx, const double *d_affected, size_t na,
                               const double *d_affecting, size_t nb, double softening, int checked,
                               double *d_out) {
    return bf::dev_call(ctx, [&] {
        return bf::dev_f64x2(ctx, d_affected, na, d_affecting, nb, softening, checked, d_out);
    });
}

int pcuda_bruteforce_f32x3_dev(pcuda_ctx *ctx, const float *d_affected, size_t na,
                               const float *d_affecting, size_t nb, float softening, int checked,
                               float *d_out) {
    return bf::dev_call(ctx, [&] {
        return bf::dev_f32x3(ctx, d_affected, na, d_affecting, nb, softening, checked, d_out);
    });
}

int pcuda_bruteforce_f32x2_dev(pcuda_ctx *ctx, const float *d_affected, size_t na,
                               const float *d_affecting, size_t nb, float softening, int checked,
                               float *d_out) {
    return bf::dev_call(ctx, [&] {
        return bf::dev_f32x2(ctx, d_affected, na, d_affecting, nb, softening, checked, d_out);
    });
}

int pcuda_bruteforce_f64x3_dev(pcuda_ctx *ctx, const double *d_affected, size_t na,
                               const double *d_affecting, size_t nb, double softening, int checked,
                               double *d_out) {
    return bf::dev_call(ctx, [&] {
        return bf::dev_f64x3(ctx, d_affected, na, d_affecting, nb, softening, checked, d_out);
    });
}

int pcuda_bruteforce_f32x3_sharded_dev(pcuda_ctx *ctx, const float *d_local_xyzm, size_t n_local,
                                       size_t shard_capacity, float softening, int checked,
                                       float *d_gathered_xyzm, float *d_out_xyz) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int s = bf::sharded_f32x3(ctx, d_local_xyzm, n_local, shard_capacity, softening, checked,
                              d_gathered_xyzm, d_out_xyz);
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_bruteforce_f32x3_sharded(pcuda_ctx *ctx, const float *local_xyzm, size_t n_local,
                                   size_t shard_capacity, float softening, int checked,
                                   float *out_xyz) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n_local && (!local_xyzm || !out_xyz))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = shard_capacity;
    if (cap == 0 || n_local > cap)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "n_local (%zu) exceeds shard capacity (%zu)",
                    n_local, cap);
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(cap * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure((size_t)world * cap * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(cap * 12));
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affecting.p, local_xyzm, n_local * 16,
                                            cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    // every rank must enter the collective, even one that owns no particles
    PCUDA_TRY(bf::sharded_f32x3(ctx, ctx->d_affecting.as<float>(), n_local, cap, softening,
                                checked, ctx->d_packed_src.as<float>(), ctx->d_out.as<float>()));
    phase_begin(ctx, PH_DOWNLOAD);
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out_xyz, ctx->d_out.p, n_local * 12,
                                            cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

int pcuda_bruteforce_f32x3_between_sharded_dev(pcuda_ctx *ctx, const float *d_affected_xyz,
                                               size_t n_affected, const float *d_local_src_xyzm,
                                               size_t n_local_src, size_t src_capacity,
                                               float softening, int checked,
                                               float *d_gathered_src_xyzm, float *d_out_xyz) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n_affected && (!d_affected_xyz || !d_out_xyz))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    size_t nb = 0;
    int s = bf::gather_sources_f32x3(ctx, d_local_src_xyzm, n_local_src, src_capacity,
                                     d_gathered_src_xyzm, &nb);
    if (s == PCUDA_OK) {
        phase_begin(ctx, PH_COMPUTE);
        s = bf::run_f32<3>(ctx, d_affected_xyz, 3, n_affected,
                           reinterpret_cast<const float4 *>(d_gathered_src_xyzm), nb, softening,
                           checked, d_out_xyz);
        phase_end(ctx, PH_COMPUTE);
    }
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_bruteforce_f32x3_between_sharded(pcuda_ctx *ctx, const float *affected_xyz,
                                           size_t n_affected, const float *local_src_xyzm,
                                           size_t n_local_src, size_t src_capacity, float softening,
                                           int checked, float *out_xyz) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if ((n_affected && (!affected_xyz || !out_xyz)) || (n_local_src && !local_src_xyzm))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = src_capacity;
    if (cap == 0 || n_local_src > cap)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "n_local (%zu) exceeds shard capacity (%zu)",
                    n_local_src, cap);
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(cap * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure((size_t)world * cap * 16));
    if (n_local_src)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affecting.p, local_src_xyzm, n_local_src * 16,
                                            cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    // every rank must enter the collective, even one that owns no targets
    size_t nb = 0;
    PCUDA_TRY(bf::gather_sources_f32x3(ctx, ctx->d_affecting.as<float>(), n_local_src, cap,
                                       ctx->d_packed_src.as<float>(), &nb));
    if (n_affected == 0) return timings_collect(ctx);
    phase_begin(ctx, PH_UPLOAD);
    return bf::host_targets<float>(ctx, affected_xyz, n_affected, 3, ctx->d_packed_src.as<float>(), nb,
                                   out_xyz, [&](float *dt, size_t n, float *ds, float *dout) {
                                       return bf::dev_f32x3(ctx, dt, n, ds, nb, softening, checked, dout);
                                   });
}

// Tuning hook: force the targets-per-thread variant (0 = automatic).  Not part of the stable ABI.
int pcuda_debug_set(const char *key, int value) {
    if (key && std::string(key) == "bf_tp") {
        bf::g_force_tp = value;
        return PCUDA_OK;
    }
    if (key && std::string(key) == "bf_waves" && value >= 1 && value <= 1024) {
        bf::g_waves = value;
        return PCUDA_OK;
    }
    if (key && std::string(key) == "bf_clamp" && value >= 0 && value <= 3) {
        bf::g_clamp_mode = value;
        return PCUDA_OK;
    }
    return pcuda::bh_debug_set(key, value);
}

}  // extern "C"
