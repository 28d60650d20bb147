// bh_build.cu — K2..K4 of the Barnes-Hut path: root cube, Morton keys, sort, gather, the level-wise
// and the single-block tree builds, quadrupoles and the double-precision layer; host-side build
// drivers.  (The one-pass build lives in bh_radix_build.cu; the traversal in bh_traverse.cu; see
// barneshut.cu for the overview of the path and the reference lines it replaces.)
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>

#include "bh.cuh"
#include "ptx.cuh"

namespace pcuda {
namespace bh {

// ------------------------------------------------------------------------------------------------
// K2a: per-axis min / max.  min/max are exact and associative, so any reduction order gives the
// bits of the sequential fold in tree/partition.rs:109-132.  NaNs are ignored (as `v < lo` does).
template <int DIM>
__global__ void __launch_bounds__(256) bbox_partial(const float *__restrict__ p, int stride, int n,
                                                    float *__restrict__ partial,
                                                    unsigned *__restrict__ mass_max_bits) {
    float lo[DIM], hi[DIM], mmax = 0.f;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        lo[k] = INFINITY;
        hi[k] = -INFINITY;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const float v = p[(size_t)i * stride + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
        mmax = fmaxf(mmax, fabsf(p[(size_t)i * stride + DIM]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mmax = fmaxf(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(mass_max_bits, __float_as_uint(mmax));
    __shared__ float s[8][2 * DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            s[w][k] = lo[k];
            s[w][DIM + k] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * DIM) {
        const bool is_hi = threadIdx.x >= DIM;
        float v = s[0][threadIdx.x];
        for (int j = 1; j < 8; ++j) v = is_hi ? fmaxf(v, s[j][threadIdx.x]) : fminf(v, s[j][threadIdx.x]);
        partial[blockIdx.x * 2 * DIM + threadIdx.x] = v;
    }
}

// K2b: final reduction + frame.  ext = max_k(hi-lo) folded from 0; half = ext/2;
// origin_k = (lo_k+hi_k)/2 - half; inv = 2^BITS/ext (0 when ext == 0).  Explicit _rn intrinsics:
// no contraction, IEEE division — the same bits as the CPU statement of the specification.
template <int DIM>
__global__ void frame_kernel(const float *__restrict__ partial, int nblocks, int n,
                             const unsigned *__restrict__ mass_max_bits, Frame *out) {
    __shared__ float s[2 * DIM];
    __shared__ float sw[8][2 * DIM];
    float v[2 * DIM];
#pragma unroll
    for (int c = 0; c < 2 * DIM; ++c) v[c] = c >= DIM ? -INFINITY : INFINITY;
    for (int j = threadIdx.x; j < nblocks; j += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 2 * DIM; ++c) {
            const float q = partial[j * 2 * DIM + c];
            v[c] = c >= DIM ? fmaxf(v[c], q) : fminf(v[c], q);
        }
    }
#pragma unroll
    for (int c = 0; c < 2 * DIM; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float q = __shfl_xor_sync(0xffffffffu, v[c], o);
            v[c] = c >= DIM ? fmaxf(v[c], q) : fminf(v[c], q);
        }
        if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5][c] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 2 * DIM) {
        const bool is_hi = threadIdx.x >= DIM;
        float r = sw[0][threadIdx.x];
        for (int j = 1; j < (int)(blockDim.x >> 5); ++j)
            r = is_hi ? fmaxf(r, sw[j][threadIdx.x]) : fminf(r, sw[j][threadIdx.x]);
        s[threadIdx.x] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ext = 0.0f;
        for (int k = 0; k < DIM; ++k) {
            const float e = __fsub_rn(s[DIM + k], s[k]);
            ext = e > ext ? e : ext;
        }
        const float half = __fdiv_rn(ext, 2.0f);
        for (int k = 0; k < 3; ++k)
            out->origin[k] = k < DIM ? __fsub_rn(__fdiv_rn(__fadd_rn(s[k], s[DIM + k]), 2.0f), half) : 0.f;
        out->ext = ext;
        out->inv = ext > 0.0f ? __fdiv_rn((float)(1ull << Dims<DIM>::BITS), ext) : 0.0f;
        out->mass_bound = (float)n * __uint_as_float(*mass_max_bits);
    }
}

__device__ __forceinline__ uint64_t spread3(uint32_t q) {  // 21 bits -> every third bit
    uint64_t x = q & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__device__ __forceinline__ uint64_t spread2(uint32_t q) {  // 31 bits -> every second bit
    uint64_t x = q & 0x7fffffffu;
    x = (x | x << 16) & 0x0000ffff0000ffffull;
    x = (x | x << 8) & 0x00ff00ff00ff00ffull;
    x = (x | x << 4) & 0x0f0f0f0f0f0f0f0full;
    x = (x | x << 2) & 0x3333333333333333ull;
    x = (x | x << 1) & 0x5555555555555555ull;
    return x;
}

template <int DIM>
__device__ __forceinline__ uint64_t encode(const float *pos, const Frame &f) {
    constexpr int BITS = Dims<DIM>::BITS;
    const float top = (float)(1ull << BITS);
    uint32_t q[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        float t = __fmul_rn(__fsub_rn(pos[k], f.origin[k]), f.inv);
        t = t > 0.0f ? t : 0.0f;  // also maps NaN to 0
        q[k] = t >= top ? (uint32_t)((1ull << BITS) - 1) : (uint32_t)t;
    }
    if (DIM == 3) return spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[DIM - 1]) << 2;
    return spread2(q[0]) | spread2(q[1]) << 1;
}

// K2c: keys in input order + identity permutation.
template <int DIM>
__global__ void __launch_bounds__(256) encode_kernel(const float *__restrict__ p, int stride, int n,
                                                     const Frame *__restrict__ frame,
                                                     uint64_t *__restrict__ keys,
                                                     uint32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Frame f = *frame;
    float pos[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) pos[k] = p[(size_t)i * stride + k];
    keys[i] = encode<DIM>(pos, f);
    idx[i] = (uint32_t)i;
}

// K3b: gather into key order as {x, y, z|0, mu}.  has_mass == false: bare positions (targets).
template <int DIM>
__global__ void __launch_bounds__(256) gather_kernel(const float *__restrict__ p, int stride,
                                                     bool has_mass, int n,
                                                     const uint32_t *__restrict__ perm,
                                                     float4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *q = p + (size_t)perm[i] * stride;
    sorted[i] = make_float4(q[0], q[1], DIM == 3 ? q[2] : 0.f, has_mass ? q[DIM] : 0.f);
}

// ------------------------------------------------------------------------------------------------
// K4: level-by-level linear orthtree WITHOUT host round trips.  The level bounds live in device
// memory (BuildState); one kernel per level is enqueued for all BITS levels up front and a kernel
// whose level turns out empty returns at once.

template <int DIM>
__device__ __forceinline__ uint32_t next_digit_start(const uint64_t *__restrict__ keys, uint32_t pos,
                                                     uint32_t end, int shift) {
    // first index in (pos, end] whose digit prefix differs from keys[pos] (keys are sorted)
    const uint64_t pre = keys[pos] >> shift;
    uint32_t lo = pos + 1, hi = end;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if ((keys[mid] >> shift) > pre) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

__global__ void init_build(NodeRec *nodes, uint32_t n, BuildState *st, uint32_t capacity) {
    NodeRec r;
    r.cm = make_float4(0.f, 0.f, 0.f, 0.f);
    r.first_child = 0;
    r.nchild_level = 0;
    r.begin = 0;
    r.count = n;
    nodes[0] = r;
    for (int i = 0; i < 36; ++i) st->level_begin[i] = i == 0 ? 0u : 1u;
    for (int i = 0; i < 34; ++i) st->ticket[i] = 0;
    st->overflow = 0;
    st->capacity = capacity;
}

constexpr int EXPAND_BLOCK = 128;

// One level: every node with more than `nleaf` particles (and above the last level) is split into
// the distinct next-level digits present in its key range (binary searches); the children of the
// level are numbered in node order — breadth-first — by a single-pass scan: tiles of 128 nodes are
// handed out by an atomic ticket, scanned in the block and chained with decoupled look-back
// (tile_state word = tag << 32 | value, tag = 4 (level + 1) + {1: tile aggregate, 2: inclusive}).
template <int DIM>
__global__ void __launch_bounds__(EXPAND_BLOCK) expand_level(NodeRec *__restrict__ nodes,
                                                             const uint64_t *__restrict__ keys,
                                                             BuildState *st,
                                                             unsigned long long *tile_state,
                                                             int level, uint32_t nleaf,
                                                             uint32_t small_level) {
    constexpr int X = Dims<DIM>::X;
    const uint32_t lvl_begin = st->level_begin[level], lvl_end = st->level_begin[level + 1];
    const uint32_t lvl_count = lvl_end - lvl_begin;
    if (lvl_count == 0 || st->overflow) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->level_begin[level + 2] = lvl_end;
        return;
    }
    // A level with few nodes (the top of the tree: huge key ranges, hardly any parallelism) is
    // latency bound, so there X threads serve one node: thread d finds where digit d starts by an
    // independent binary search, instead of one thread walking from digit to digit.
    const bool small = lvl_count <= small_level;
    const uint32_t tile_nodes = small ? EXPAND_BLOCK / X : EXPAND_BLOCK;
    const uint32_t n_tiles = (lvl_count + tile_nodes - 1) / tile_nodes;
    const uint32_t capacity = st->capacity;
    const int shift = DIM * (Dims<DIM>::BITS - level - 1);
    typedef cub::BlockScan<uint32_t, EXPAND_BLOCK> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_tile, s_prefix;
    __shared__ uint32_t s_b[EXPAND_BLOCK / X][X + 1];
    const unsigned long long tag = (unsigned long long)(level + 1) * 4;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&st->ticket[level], 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        if (small) {
            const uint32_t tn = tile * tile_nodes + threadIdx.x / X;
            const uint32_t d = threadIdx.x % X;
            if (tn < lvl_count) {
                const uint32_t begin = nodes[lvl_begin + tn].begin, count = nodes[lvl_begin + tn].count;
                if (count > nleaf && level < Dims<DIM>::BITS) {
                    const uint64_t want = ((keys[begin] >> shift) & ~(uint64_t)(X - 1)) | d;
                    uint32_t lo = begin, hi = begin + count;
                    while (d != 0 && lo < hi) {  // first key of the cell whose digit is >= d
                        const uint32_t mid = lo + ((hi - lo) >> 1);
                        if ((keys[mid] >> shift) < want) lo = mid + 1;
                        else hi = mid;
                    }
                    s_b[threadIdx.x / X][d] = lo;
                    if (d == 0) s_b[threadIdx.x / X][X] = begin + count;
                }
            }
            __syncthreads();
        }
        const uint32_t t = tile * tile_nodes + threadIdx.x;
        const bool valid = threadIdx.x < tile_nodes && t < lvl_count;
        uint32_t c = 0, cb[X + 1];
        if (valid) {
            const uint32_t begin = nodes[lvl_begin + t].begin, count = nodes[lvl_begin + t].count;
            if (count > nleaf && level < Dims<DIM>::BITS) {
                const uint32_t end = begin + count;
                if (small) {
#pragma unroll
                    for (int k = 0; k < X; ++k) {
                        const uint32_t bk = s_b[threadIdx.x][k];
                        const bool present = s_b[threadIdx.x][k + 1] > bk;
#pragma unroll
                        for (int m = 0; m < X; ++m)
                            if (present && m == (int)c) cb[m] = bk;
                        c += present;
                    }
                } else {
                    uint32_t pos = begin;
#pragma unroll
                    for (int k = 0; k < X; ++k) {
                        if (pos < end) {
                            cb[k] = pos;
                            pos = next_digit_start<DIM>(keys, pos, end, shift);
                            ++c;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k <= X; ++k)
                    if (k == (int)c) cb[k] = end;
            }
        }
        uint32_t off, total;
        Scan(scan_tmp).ExclusiveSum(c, off, total);
        if (threadIdx.x == 0) {
            uint32_t excl = 0;
            volatile unsigned long long *ts = tile_state;
            if (tile > 0) {
                ts[tile] = (tag + 1) << 32 | total;
                __threadfence();
                int p = (int)tile - 1;
                for (;;) {
                    const unsigned long long w = ts[p];
                    const unsigned long long wt = w >> 32;
                    if (wt == tag + 2) {
                        excl += (uint32_t)w;
                        break;
                    }
                    if (wt == tag + 1) {
                        excl += (uint32_t)w;
                        --p;
                    }
                }
            }
            ts[tile] = (tag + 2) << 32 | (excl + total);
            __threadfence();
            s_prefix = excl;
            if (tile == n_tiles - 1) {
                const unsigned long long next_end = (unsigned long long)lvl_end + excl + total;
                if (next_end > capacity) {
                    st->overflow = 1;
                    st->level_begin[level + 2] = lvl_end;
                } else {
                    st->level_begin[level + 2] = (uint32_t)next_end;
                }
            }
        }
        __syncthreads();
        if (valid) {
            NodeRec &nd = nodes[lvl_begin + t];
            const unsigned long long first = (unsigned long long)lvl_end + s_prefix + off;
            if (c == 0 || first + c > capacity) {
                nd.first_child = 0;
                nd.nchild_level = (uint32_t)level << 8;
            } else {
                nd.first_child = (uint32_t)first;
                nd.nchild_level = c | (uint32_t)level << 8;
#pragma unroll
                for (int k = 0; k < X; ++k) {
                    if (k < (int)c) {
                        NodeRec ch;
                        ch.cm = make_float4(0.f, 0.f, 0.f, 0.f);
                        ch.first_child = 0;
                        ch.nchild_level = (uint32_t)(level + 1) << 8;
                        ch.begin = cb[k];
                        ch.count = cb[k + 1] - cb[k];
                        nodes[first + k] = ch;
                    }
                }
            }
        }
    }
}

// K4c: moments of one level, deepest level first.  Double precision, fixed order, unfused
// (__dmul_rn / __dadd_rn), identical to the CPU statement of the specification:
//   leaf:      M = sum m_i, Mx_k = sum m_i * x_ik over the cell's particles in key order
//   internal:  sums of the children's moments in child order
//   com_k = (float)(Mx_k / M), mass = (float)M;  M == 0 => com = position of the first particle.
template <int DIM>
__device__ __forceinline__ void node_moments(NodeRec *__restrict__ nodes, double *__restrict__ mom,
                                             const float4 *__restrict__ sorted, uint32_t j) {
    NodeRec nd = nodes[j];
    const uint32_t nc = nd.nchild_level & 0xffu;
    double m[4] = {0.0, 0.0, 0.0, 0.0};  // x, y, z, M
    if (nc == 0) {
        for (uint32_t i = nd.begin; i < nd.begin + nd.count; ++i) {
            const float4 p = sorted[i];
            const double mi = (double)p.w;
            m[0] = __dadd_rn(m[0], __dmul_rn(mi, (double)p.x));
            m[1] = __dadd_rn(m[1], __dmul_rn(mi, (double)p.y));
            if (DIM == 3) m[2] = __dadd_rn(m[2], __dmul_rn(mi, (double)p.z));
            m[3] = __dadd_rn(m[3], mi);
        }
    } else {
        for (uint32_t c = nd.first_child; c < nd.first_child + nc; ++c) {
            const double4 q = reinterpret_cast<const double4 *>(mom)[c];
            m[0] = __dadd_rn(m[0], q.x);
            m[1] = __dadd_rn(m[1], q.y);
            if (DIM == 3) m[2] = __dadd_rn(m[2], q.z);
            m[3] = __dadd_rn(m[3], q.w);
        }
    }
    reinterpret_cast<double4 *>(mom)[j] = make_double4(m[0], m[1], m[2], m[3]);
    float4 cm;
    if (m[3] == 0.0) {
        const float4 p = sorted[nd.begin];
        cm = make_float4(p.x, p.y, DIM == 3 ? p.z : 0.f, 0.f);
    } else {
        cm.x = (float)__ddiv_rn(m[0], m[3]);
        cm.y = (float)__ddiv_rn(m[1], m[3]);
        cm.z = DIM == 3 ? (float)__ddiv_rn(m[2], m[3]) : 0.f;
        cm.w = (float)m[3];
    }
    nodes[j].cm = cm;
}

template <int DIM>
__global__ void __launch_bounds__(128) moments_kernel(NodeRec *__restrict__ nodes,
                                                      double *__restrict__ mom,
                                                      const float4 *__restrict__ sorted,
                                                      const BuildState *__restrict__ st, int level) {
    const uint32_t lvl_begin = st->level_begin[level];
    const uint32_t lvl_count = st->level_begin[level + 1] - lvl_begin;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < lvl_count;
         t += gridDim.x * blockDim.x)
        node_moments<DIM>(nodes, mom, sorted, lvl_begin + t);
}

// K4 for small inputs: the whole tree — every level of the expansion, then every level of the
// moments — in ONE single-block launch with __syncthreads() between levels.  Below ~32k particles
// the per-level kernels above are pure launch latency (44 launches ~ 150 us for a tree that takes
// a few microseconds to build), and the reference's own benchmark lives at those sizes
// (benches/benchmark.rs: N = 2 .. 65536).  Same numbering (children in node order, breadth-first)
// and the same arithmetic as the per-level path: the arrays are bit-identical.
constexpr int SMALL_TREE_BLOCK = 1024;
constexpr size_t SMALL_TREE_MAX_N = 32768;

template <int DIM>
__global__ void __launch_bounds__(SMALL_TREE_BLOCK) build_small(NodeRec *__restrict__ nodes,
                                                                double *__restrict__ mom,
                                                                const uint64_t *__restrict__ keys,
                                                                const float4 *__restrict__ sorted,
                                                                BuildState *st, uint32_t n,
                                                                uint32_t capacity, uint32_t nleaf) {
    constexpr int X = Dims<DIM>::X;
    constexpr int BITS = Dims<DIM>::BITS;
    typedef cub::BlockScan<uint32_t, SMALL_TREE_BLOCK> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_begin[36];
    __shared__ uint32_t s_overflow;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        NodeRec r;
        r.cm = make_float4(0.f, 0.f, 0.f, 0.f);
        r.first_child = 0;
        r.nchild_level = 0;
        r.begin = 0;
        r.count = n;
        nodes[0] = r;
        for (int i = 0; i < 36; ++i) s_begin[i] = i == 0 ? 0u : 1u;
        s_overflow = 0;
    }
    __syncthreads();
    int levels = 0;
    for (int level = 0; level <= BITS; ++level) {
        const uint32_t lvl_begin = s_begin[level], lvl_end = s_begin[level + 1];
        const uint32_t lvl_count = lvl_end - lvl_begin;
        if (lvl_count == 0 || s_overflow) break;
        levels = level + 1;
        const int shift = DIM * (BITS - level - 1);
        uint32_t running = 0;  // children emitted so far on this level (uniform)
        for (uint32_t base = 0; base < lvl_count; base += SMALL_TREE_BLOCK) {
            const uint32_t t = base + tid;
            uint32_t c = 0, cb[X + 1];
            if (t < lvl_count) {
                const uint32_t begin = nodes[lvl_begin + t].begin, count = nodes[lvl_begin + t].count;
                if (count > nleaf && level < BITS) {
                    uint32_t pos = begin;
                    const uint32_t end = begin + count;
#pragma unroll
                    for (int k = 0; k < X; ++k) {
                        if (pos < end) {
                            cb[k] = pos;
                            pos = next_digit_start<DIM>(keys, pos, end, shift);
                            ++c;
                        }
                    }
#pragma unroll
                    for (int k = 0; k <= X; ++k)
                        if (k == (int)c) cb[k] = end;
                }
            }
            uint32_t off, total;
            __syncthreads();  // scan_tmp reuse
            Scan(scan_tmp).ExclusiveSum(c, off, total);
            const unsigned long long first = (unsigned long long)lvl_end + running + off;
            if ((unsigned long long)lvl_end + running + total > capacity) {
                if (tid == 0) s_overflow = 1;
                c = 0;
            }
            if (t < lvl_count) {
                NodeRec &nd = nodes[lvl_begin + t];
                if (c == 0) {
                    nd.first_child = 0;
                    nd.nchild_level = (uint32_t)level << 8;
                } else {
                    nd.first_child = (uint32_t)first;
                    nd.nchild_level = c | (uint32_t)level << 8;
#pragma unroll
                    for (int k = 0; k < X; ++k) {
                        if (k < (int)c) {
                            NodeRec ch;
                            ch.cm = make_float4(0.f, 0.f, 0.f, 0.f);
                            ch.first_child = 0;
                            ch.nchild_level = (uint32_t)(level + 1) << 8;
                            ch.begin = cb[k];
                            ch.count = cb[k + 1] - cb[k];
                            nodes[first + k] = ch;
                        }
                    }
                }
            }
            running += total;
        }
        __syncthreads();
        if (tid == 0 && !s_overflow) s_begin[level + 2] = lvl_end + running;
        __syncthreads();
    }
    __syncthreads();  // every thread has read the level table of the iteration that left the loop
    if (tid == 0) {
        // levels past the last one are empty: level_begin stays at the end of the last level
        for (int l = levels + 1; l < 36; ++l) s_begin[l] = s_begin[levels];
    }
    __syncthreads();
    if (!s_overflow) {
        for (int level = levels - 1; level >= 0; --level) {
            const uint32_t lvl_begin = s_begin[level], lvl_count = s_begin[level + 1] - lvl_begin;
            for (uint32_t t = tid; t < lvl_count; t += SMALL_TREE_BLOCK)
                node_moments<DIM>(nodes, mom, sorted, lvl_begin + t);
            __syncthreads();
        }
    }
    if (tid < 36) st->level_begin[tid] = s_begin[tid];
    if (tid < 34) st->ticket[tid] = 0;  // unused here; the host reads the whole state back
    if (tid == 0) {
        st->overflow = s_overflow;
        st->capacity = capacity;
    }
}

// ------------------------------------------------------------------------------------------------
// K4d / K5q: quadrupole nodes (pcuda_config.expansion_order = 2; beyond the reference, whose nodes
// carry {centre of mass, mass} only, gravity/impls/mod.rs:103-135).  Every node additionally holds
// the traceless quadrupole about its centre of mass,
//     Q = sum_i m_i (3 x_i x_i^T - |x_i|^2 I),   x_i = p_i - com,
// built bottom-up in double precision (leaves from their particles, internal nodes from their
// children with the parallel-axis term m_c (3 d d^T - |d|^2 I), d = com_c - com).  An accepted
// node then contributes, with D = com - target and R = |D|,
//     a = M D / R^3  -  Q D / R^5  +  5/2 (D.Q.D) D / R^7,
// evaluated as  ri^2 [ (M + 5/2 u.Qu') u - Qu' ],  u = D ri,  Qu' = (Q u) ri^2, so that no
// intermediate exceeds the magnitude of the monopole term's own factors.
template <int DIM>
__global__ void __launch_bounds__(128) quad_kernel(const NodeRec *__restrict__ nodes,
                                                   const double4 *__restrict__ mom,
                                                   const float4 *__restrict__ sorted,
                                                   double *__restrict__ quad64,
                                                   float4 *__restrict__ quadf,
                                                   const BuildState *__restrict__ st, int level) {
    const uint32_t lvl_begin = st->level_begin[level];
    const uint32_t lvl_count = st->level_begin[level + 1] - lvl_begin;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < lvl_count;
         t += gridDim.x * blockDim.x) {
        const uint32_t j = lvl_begin + t;
        const NodeRec nd = nodes[j];
        const uint32_t nc = nd.nchild_level & 0xffu;
        const double4 sm = mom[j];
        double q[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // xx xy xz yy yz zz
        if (sm.w != 0.0) {
            const double cx = sm.x / sm.w, cy = sm.y / sm.w, cz = DIM == 3 ? sm.z / sm.w : 0.0;
            auto add = [&](double m, double x, double y, double z) {
                const double r2 = x * x + y * y + z * z;
                q[0] += m * (3.0 * x * x - r2);
                q[1] += m * (3.0 * x * y);
                q[2] += m * (3.0 * x * z);
                q[3] += m * (3.0 * y * y - r2);
                q[4] += m * (3.0 * y * z);
                q[5] += m * (3.0 * z * z - r2);
            };
            if (nc == 0) {
                for (uint32_t i = nd.begin; i < nd.begin + nd.count; ++i) {
                    const float4 p = sorted[i];
                    add((double)p.w, (double)p.x - cx, (double)p.y - cy, DIM == 3 ? (double)p.z - cz : 0.0);
                }
            } else {
                for (uint32_t c = nd.first_child; c < nd.first_child + nc; ++c) {
                    const double4 sc = mom[c];
                    if (sc.w == 0.0) continue;
                    add(sc.w, sc.x / sc.w - cx, sc.y / sc.w - cy, DIM == 3 ? sc.z / sc.w - cz : 0.0);
#pragma unroll
                    for (int k = 0; k < 6; ++k) q[k] += quad64[(size_t)c * 6 + k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) quad64[(size_t)j * 6 + k] = q[k];
        quadf[2 * (size_t)j] = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
        quadf[2 * (size_t)j + 1] = make_float4((float)q[4], (float)q[5], 0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// K5d: double precision (DVec2 / DVec3 particles; the reference's BarnesHut is generic over the
// scalar, sequential.rs:439-543).  The TREE STRUCTURE — keys, sort, cells, opening decisions — is
// the f32 one, built over the particles rounded to f32 (an opening decision moved by 2^-24 of the
// box size is immaterial).  Everything that enters an acceleration is double precision: the
// sources in key order (double4), the centre of mass of every node (recomputed bottom-up from the
// f64 positions), the targets, and the pair term (the 16-operation FP64 sequence of the f64
// brute-force kernel).  theta = 0 opens every cell, so the result is the f64 brute-force sum.
__global__ void __launch_bounds__(256) narrow_kernel(const double *__restrict__ in, size_t count,
                                                     float *__restrict__ out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count;
         i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i];
}

template <int DIM>
__global__ void __launch_bounds__(256) gather64_kernel(const double *__restrict__ p, int stride,
                                                       bool has_mass, int n,
                                                       const uint32_t *__restrict__ perm,
                                                       double4 *__restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *q = p + (size_t)perm[i] * stride;
    sorted[i] = make_double4(q[0], q[1], DIM == 3 ? q[2] : 0.0, has_mass ? q[DIM] : 0.0);
}

// Bottom-up sums {sum m x, sum m y, sum m z, sum m} of one level from the f64 records (leaves) or
// the children's sums (internal nodes), same fixed order as node_moments.
template <int DIM>
__global__ void __launch_bounds__(128) moments64_kernel(const NodeRec *__restrict__ nodes,
                                                        double4 *__restrict__ mom,
                                                        const double4 *__restrict__ sorted64,
                                                        const BuildState *__restrict__ st, int level) {
    const uint32_t lvl_begin = st->level_begin[level];
    const uint32_t lvl_count = st->level_begin[level + 1] - lvl_begin;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < lvl_count;
         t += gridDim.x * blockDim.x) {
        const uint32_t j = lvl_begin + t;
        const NodeRec nd = nodes[j];
        const uint32_t nc = nd.nchild_level & 0xffu;
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        if (nc == 0) {
            for (uint32_t i = nd.begin; i < nd.begin + nd.count; ++i) {
                const double4 q = sorted64[i];
                m[0] = __dadd_rn(m[0], __dmul_rn(q.w, q.x));
                m[1] = __dadd_rn(m[1], __dmul_rn(q.w, q.y));
                if (DIM == 3) m[2] = __dadd_rn(m[2], __dmul_rn(q.w, q.z));
                m[3] = __dadd_rn(m[3], q.w);
            }
        } else {
            for (uint32_t c = nd.first_child; c < nd.first_child + nc; ++c) {
                const double4 q = mom[c];
                m[0] = __dadd_rn(m[0], q.x);
                m[1] = __dadd_rn(m[1], q.y);
                if (DIM == 3) m[2] = __dadd_rn(m[2], q.z);
                m[3] = __dadd_rn(m[3], q.w);
            }
        }
        mom[j] = make_double4(m[0], m[1], m[2], m[3]);
    }
}

// sums -> {com, mass} in place (a massless cell sits at its first particle, as in the f32 tree).
__global__ void __launch_bounds__(256) finalize_cm64(const NodeRec *__restrict__ nodes,
                                                     double4 *__restrict__ mom,
                                                     const double4 *__restrict__ sorted64,
                                                     uint32_t n_nodes) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_nodes; j += gridDim.x * blockDim.x) {
        const double4 q = mom[j];
        if (q.w == 0.0) {
            const double4 f = sorted64[nodes[j].begin];
            mom[j] = make_double4(f.x, f.y, f.z, 0.0);
        } else {
            mom[j] = make_double4(__ddiv_rn(q.x, q.w), __ddiv_rn(q.y, q.w), __ddiv_rn(q.z, q.w), q.w);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side.
template <int DIM>
int sort_by_key(pcuda_ctx *ctx, const float *d_pos, int stride, size_t n, const Frame *d_frame,
                       DevBuf keys[2], DevBuf perm[2], int *cur, DevBuf &cub_tmp) {
    for (int i = 0; i < 2; ++i) {
        PCUDA_CUDA_TRY(ctx, keys[i].ensure(n * sizeof(uint64_t)));
        PCUDA_CUDA_TRY(ctx, perm[i].ensure(n * sizeof(uint32_t)));
    }
    encode_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        d_pos, stride, (int)n, d_frame, keys[0].as<uint64_t>(), perm[0].as<uint32_t>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    cub::DoubleBuffer<uint64_t> kb(keys[0].as<uint64_t>(), keys[1].as<uint64_t>());
    cub::DoubleBuffer<uint32_t> vb(perm[0].as<uint32_t>(), perm[1].as<uint32_t>());
    size_t tmp = 0;
    const int end_bit = DIM * Dims<DIM>::BITS;
    PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int)n, 0, end_bit,
                                                        ctx->stream));
    PCUDA_CUDA_TRY(ctx, cub_tmp.ensure(tmp));
    PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp, kb, vb, (int)n, 0, end_bit,
                                                        ctx->stream));
    ctx->launches += 1 + (end_bit + 7) / 8;  // histogram + one onesweep pass per 8 bits
    *cur = kb.selector;
    return PCUDA_OK;
}

// Resets the host-side description of `t` for a tree of `n` particles.
template <int DIM>
void tree_reset(pcuda_ctx *ctx, pcuda_tree *t, size_t n) {
    t->dim = DIM;
    t->bits = Dims<DIM>::BITS;
    t->n = n;
    t->n_nodes = 0;
    t->n_levels = 0;
    t->leaf_size = ctx->leaf_size;
    t->level_begin.clear();
    t->frame = Frame{};
    t->d_parent = nullptr;
}

// K2: root cube of `n` particle rows -> t->d_frame (device).
template <int DIM>
int build_frame(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n) {
    const int stride = DIM + 1;
    cudaStream_t st = ctx->stream;
    const int nb = (int)std::min<size_t>(ctx->sm_count * 8, (n + 255) / 256);
    PCUDA_CUDA_TRY(ctx, t->partial.ensure((size_t)nb * 2 * DIM * sizeof(float)));
    PCUDA_CUDA_TRY(ctx, t->d_frame.ensure(sizeof(Frame) + sizeof(unsigned)));
    unsigned *d_mmax = reinterpret_cast<unsigned *>(t->d_frame.as<Frame>() + 1);
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_mmax, 0, sizeof(unsigned), st));
    bbox_partial<DIM><<<nb, 256, 0, st>>>(d_particles, stride, (int)n, t->partial.as<float>(), d_mmax);
    frame_kernel<DIM><<<1, 256, 0, st>>>(t->partial.as<float>(), nb, (int)n, d_mmax,
                                        t->d_frame.as<Frame>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PCUDA_OK;
}

template <int DIM>
int build(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n, bool keys_only) {
    const int stride = DIM + 1;
    tree_reset<DIM>(ctx, t, n);
    if (n == 0) return PCUDA_OK;
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    cudaStream_t st = ctx->stream;

    // K2: root cube + keys
    PCUDA_TRY(build_frame<DIM>(ctx, t, d_particles, n));
    // K3: sort + gather
    PCUDA_TRY(sort_by_key<DIM>(ctx, d_particles, stride, n, t->d_frame.as<Frame>(), t->keys, t->perm,
                               &t->cur, t->cub_tmp));
    if (keys_only) {  // pcuda_morton_*: root cube, keys and sort permutation only
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(&t->frame, t->d_frame.p, sizeof(Frame), cudaMemcpyDeviceToHost, st));
        PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        return PCUDA_OK;
    }
    PCUDA_CUDA_TRY(ctx, t->sorted.ensure(n * sizeof(float4)));
    gather_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        d_particles, stride, true, (int)n, t->d_perm(), t->sorted.as<float4>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return build_levels<DIM>(ctx, t, n);
}

// K4: level-by-level linear orthtree over t->d_keys() / t->sorted (n sorted particles, frame in
// t->d_frame), all levels enqueued without host round trips; one read-back of the level table at
// the end.  If the node capacity guess was too small the build is repeated with the capacity it
// asked for (grow-only, so this happens at most once per size class).
template <int DIM>
int build_levels(pcuda_ctx *ctx, pcuda_tree *t, size_t n) {
    constexpr int BITS = Dims<DIM>::BITS;
    cudaStream_t st = ctx->stream;
    for (int attempt = 0;; ++attempt) {
        size_t cap_nodes = std::max<size_t>(4096, (size_t)((double)n * t->nodes_per_particle) + 1024);
        PCUDA_CUDA_TRY(ctx, t->nodes.ensure(cap_nodes * sizeof(NodeRec)));
        cap_nodes = std::min<size_t>(t->nodes.cap / sizeof(NodeRec), 0xfffffff0ull);
        PCUDA_CUDA_TRY(ctx, t->moments.ensure(cap_nodes * 4 * sizeof(double)));
        // tiles of 128 nodes, or of 128 / 2^DIM nodes on levels of <= SMALL_LEVEL nodes
        const size_t max_tiles = cap_nodes / (EXPAND_BLOCK / Dims<DIM>::X) + 2;
        PCUDA_CUDA_TRY(ctx, t->scan_in.ensure(sizeof(BuildState)));
        PCUDA_CUDA_TRY(ctx, t->scan_out.ensure(max_tiles * sizeof(unsigned long long)));
        BuildState *d_state = t->scan_in.as<BuildState>();
        if (n <= SMALL_TREE_MAX_N && g_level_build != 2) {
            build_small<DIM><<<1, SMALL_TREE_BLOCK, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double>(),
                                                             t->d_keys(), t->sorted.as<float4>(), d_state,
                                                             (uint32_t)n, (uint32_t)cap_nodes, t->leaf_size);
            PCUDA_CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches += 1;
        } else if (g_level_build != 1 && t->leaf_size <= (uint32_t)RB_MAX_LEAF) {
            PCUDA_TRY(radix_build_enqueue<DIM>(ctx, t, n, cap_nodes, d_state));
        } else {  // level-wise build (tuning hook bh_level_build; leaf sizes beyond the one-pass window)
            PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(t->scan_out.p, 0, max_tiles * sizeof(unsigned long long), st));
            init_build<<<1, 1, 0, st>>>(t->nodes.as<NodeRec>(), (uint32_t)n, d_state, (uint32_t)cap_nodes);
            const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, max_tiles);
            for (int level = 0; level <= BITS; ++level)
                expand_level<DIM><<<grid, EXPAND_BLOCK, 0, st>>>(
                    t->nodes.as<NodeRec>(), t->d_keys(), d_state,
                    t->scan_out.as<unsigned long long>(), level, t->leaf_size, g_small_level);
            const unsigned mgrid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, (cap_nodes + 127) / 128);
            for (int level = BITS; level >= 0; --level)
                moments_kernel<DIM><<<mgrid, 128, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double>(),
                                                           t->sorted.as<float4>(), d_state, level);
            PCUDA_CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches += 1 + 2 * (BITS + 1);
        }
        BuildState h;
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(&t->frame, t->d_frame.p, sizeof(Frame), cudaMemcpyDeviceToHost, st));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(&h, d_state, sizeof h, cudaMemcpyDeviceToHost, st));
        PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        if (h.overflow) {
            if (attempt >= 8 || cap_nodes >= 0xfffffff0ull)
                return fail(ctx, PCUDA_ERR_TREE_OVERFLOW, "tree does not fit into %zu nodes", cap_nodes);
            t->nodes_per_particle = std::max(2.0 * t->nodes_per_particle, 2.0 * (double)cap_nodes / (double)n);
            continue;
        }
        t->level_begin.clear();
        int levels = 0;
        while (levels <= BITS && h.level_begin[levels + 1] > h.level_begin[levels]) ++levels;
        for (int l = 0; l <= levels; ++l) t->level_begin.push_back(h.level_begin[l]);
        t->n_levels = levels;
        t->n_nodes = h.level_begin[levels];
        break;
    }
    t->order = (int)ctx->order;
    if (t->order == 2) {  // quadrupoles, bottom-up (K4d)
        PCUDA_CUDA_TRY(ctx, t->quad64.ensure(t->n_nodes * 6 * sizeof(double)));
        PCUDA_CUDA_TRY(ctx, t->quad.ensure(t->n_nodes * 2 * sizeof(float4)));
        const BuildState *d_state = t->scan_in.as<BuildState>();
        for (int level = t->n_levels - 1; level >= 0; --level) {
            const uint32_t cnt = t->level_begin[level + 1] - t->level_begin[level];
            const unsigned grid = std::min<unsigned>((unsigned)ctx->sm_count * 8, (cnt + 127) / 128);
            quad_kernel<DIM><<<grid, 128, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double4>(),
                                                   t->sorted.as<float4>(), t->quad64.as<double>(),
                                                   t->quad.as<float4>(), d_state, level);
        }
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += t->n_levels;
    }
    return PCUDA_OK;
}

int build_dim(pcuda_ctx *ctx, pcuda_tree *t, uint32_t dim, const float *d_particles, size_t n,
                     bool keys_only) {
    if (dim == 3) return build<3>(ctx, t, d_particles, n, keys_only);
    if (dim == 2) return build<2>(ctx, t, d_particles, n, keys_only);
    return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
}

// Tree over double-precision particles: the f32 structure over the rounded records, then the f64
// layer (sources in key order, {com, mass} per node from the f64 positions, bottom-up).
template <int DIM>
int build64(pcuda_ctx *ctx, pcuda_tree *t, const double *d_particles64, size_t n) {
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    cudaStream_t st = ctx->stream;
    const size_t count = n * (DIM + 1);
    if (n) {
        PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure(count * sizeof(float)));
        narrow_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 65535), 256, 0, st>>>(
            d_particles64, count, ctx->d_packed_src.as<float>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    PCUDA_TRY(build<DIM>(ctx, t, ctx->d_packed_src.as<float>(), n));
    if (n == 0) return PCUDA_OK;
    PCUDA_CUDA_TRY(ctx, t->sorted64.ensure(n * sizeof(double4)));
    gather64_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        d_particles64, DIM + 1, true, (int)n, t->d_perm(), t->sorted64.as<double4>());
    const BuildState *d_state = t->scan_in.as<BuildState>();
    for (int level = t->n_levels - 1; level >= 0; --level) {
        const uint32_t cnt = t->level_begin[level + 1] - t->level_begin[level];
        const unsigned grid = std::min<unsigned>((unsigned)ctx->sm_count * 8, (cnt + 127) / 128);
        moments64_kernel<DIM><<<grid, 128, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double4>(),
                                                    t->sorted64.as<double4>(), d_state, level);
    }
    finalize_cm64<<<std::min<unsigned>((unsigned)ctx->sm_count * 8, (unsigned)((t->n_nodes + 255) / 256)), 256, 0,
                    st>>>(t->nodes.as<NodeRec>(), t->moments.as<double4>(), t->sorted64.as<double4>(),
                          (uint32_t)t->n_nodes);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2 + t->n_levels;
    return PCUDA_OK;
}


// ---- root cube of a cloud that is spread over several ranks (bh_multigpu.cu) -------------------------
// Every rank reduces its own records to {lo[3], hi[3], max|mu|, 0} (local_box); the boxes of all
// ranks are exchanged; frame_from_boxes folds them — min / max are exact and associative, so the
// frame has the bits of the single-GPU frame over all particles.
__global__ void box_kernel(const float *__restrict__ partial, int nblocks, const unsigned *__restrict__ mass_max_bits,
                           float *__restrict__ box8) {
    if (threadIdx.x < 6) {
        const bool is_hi = threadIdx.x >= 3;
        float v = is_hi ? -INFINITY : INFINITY;
        for (int j = 0; j < nblocks; ++j) {
            const float q = partial[j * 6 + threadIdx.x];
            v = is_hi ? fmaxf(v, q) : fminf(v, q);
        }
        box8[threadIdx.x] = v;
    }
    if (threadIdx.x == 6) box8[6] = __uint_as_float(*mass_max_bits);
    if (threadIdx.x == 7) box8[7] = 0.f;
}

__global__ void frame_from_boxes_kernel(const float *__restrict__ boxes, int world, unsigned long long n_total,
                                        Frame *out) {
    if (threadIdx.x != 0) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, mmax = 0.f;
    for (int r = 0; r < world; ++r) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], boxes[r * 8 + k]);
            hi[k] = fmaxf(hi[k], boxes[r * 8 + 3 + k]);
        }
        mmax = fmaxf(mmax, boxes[r * 8 + 6]);
    }
    float ext = 0.0f;  // same arithmetic as frame_kernel
    for (int k = 0; k < 3; ++k) {
        const float e = __fsub_rn(hi[k], lo[k]);
        ext = e > ext ? e : ext;
    }
    const float half = __fdiv_rn(ext, 2.0f);
    for (int k = 0; k < 3; ++k) out->origin[k] = __fsub_rn(__fdiv_rn(__fadd_rn(lo[k], hi[k]), 2.0f), half);
    out->ext = ext;
    out->inv = ext > 0.0f ? __fdiv_rn((float)(1ull << Dims<3>::BITS), ext) : 0.0f;
    out->mass_bound = (float)n_total * mmax;
}

int local_box(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n, float *d_box8) {
    cudaStream_t st = ctx->stream;
    const int nb = (int)std::max<size_t>(1, std::min<size_t>(ctx->sm_count * 8, (n + 255) / 256));
    PCUDA_CUDA_TRY(ctx, t->partial.ensure((size_t)nb * 6 * sizeof(float)));
    PCUDA_CUDA_TRY(ctx, t->d_frame.ensure(sizeof(Frame) + sizeof(unsigned)));
    unsigned *d_mmax = reinterpret_cast<unsigned *>(t->d_frame.as<Frame>() + 1);
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_mmax, 0, sizeof(unsigned), st));
    bbox_partial<3><<<nb, 256, 0, st>>>(d_particles, 4, (int)n, t->partial.as<float>(), d_mmax);  // n == 0: +-inf
    box_kernel<<<1, 32, 0, st>>>(t->partial.as<float>(), nb, d_mmax, d_box8);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PCUDA_OK;
}

int frame_from_boxes(pcuda_ctx *ctx, pcuda_tree *t, const float *d_boxes, int world, size_t n_total) {
    PCUDA_CUDA_TRY(ctx, t->d_frame.ensure(sizeof(Frame) + sizeof(unsigned)));
    frame_from_boxes_kernel<<<1, 32, 0, ctx->stream>>>(d_boxes, world, (unsigned long long)n_total, t->d_frame.as<Frame>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return PCUDA_OK;
}

// Launchers for the kernels the other translation units need (the kernels themselves stay here).
template <int DIM>
void launch_encode(pcuda_ctx *ctx, const float *d_pos, int stride, size_t n, const Frame *d_frame,
                   uint64_t *keys, uint32_t *idx) {
    encode_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_pos, stride, (int)n, d_frame, keys, idx);
    ctx->launches++;
}

template <int DIM>
void launch_gather(pcuda_ctx *ctx, const float *d_pos, int stride, bool has_mass, size_t n,
                   const uint32_t *perm, float4 *sorted) {
    gather_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_pos, stride, has_mass, (int)n, perm, sorted);
    ctx->launches++;
}

template <int DIM>
void launch_gather64(pcuda_ctx *ctx, const double *d_pos, int stride, bool has_mass, size_t n,
                     const uint32_t *perm, double4 *sorted) {
    gather64_kernel<DIM><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_pos, stride, has_mass, (int)n, perm, sorted);
    ctx->launches++;
}

void launch_narrow(pcuda_ctx *ctx, const double *in, size_t count, float *out) {
    narrow_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 65535), 256, 0, ctx->stream>>>(in, count, out);
    ctx->launches++;
}

#define PCUDA_BH_INSTANTIATE(DIM)                                                                              \
    template int sort_by_key<DIM>(pcuda_ctx *, const float *, int, size_t, const Frame *, DevBuf[2], DevBuf[2], \
                                  int *, DevBuf &);                                                            \
    template void tree_reset<DIM>(pcuda_ctx *, pcuda_tree *, size_t);                                          \
    template int build_frame<DIM>(pcuda_ctx *, pcuda_tree *, const float *, size_t);                           \
    template int build_levels<DIM>(pcuda_ctx *, pcuda_tree *, size_t);                                         \
    template int build<DIM>(pcuda_ctx *, pcuda_tree *, const float *, size_t, bool);                           \
    template int build64<DIM>(pcuda_ctx *, pcuda_tree *, const double *, size_t);                              \
    template void launch_encode<DIM>(pcuda_ctx *, const float *, int, size_t, const Frame *, uint64_t *,       \
                                     uint32_t *);                                                              \
    template void launch_gather<DIM>(pcuda_ctx *, const float *, int, bool, size_t, const uint32_t *, float4 *); \
    template void launch_gather64<DIM>(pcuda_ctx *, const double *, int, bool, size_t, const uint32_t *, double4 *);
PCUDA_BH_INSTANTIATE(2)
PCUDA_BH_INSTANTIATE(3)
#undef PCUDA_BH_INSTANTIATE

}  // namespace bh
}  // namespace pcuda
