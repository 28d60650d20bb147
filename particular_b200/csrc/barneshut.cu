// barneshut.cu — K2..K5: Barnes-Hut on the GPU (new: the reference has only CPU versions,
// particular/src/sequential.rs:439-543 and parallel.rs:297-367, whose tree is the recursive
// bucket partition of tree/mod.rs:91-138).
//
// Build (K2-K4), every call, like the reference (sequential.rs:539-541):
//   K2  root cube = BoundingBox::square_with (tree/partition.rs:136-153): min/max reduction, then
//       origin/extent/scale;  Morton (3-D, 21 bits/axis) / Z-order (2-D, 31 bits/axis) keys
//   K3  stable LSD radix sort of (key, index) pairs (cub::DeviceRadixSort), gather of the
//       particles into key order as float4 {x, y, z|0, mu}
//   K4  linear orthtree over the sorted keys, level by level (a cell with more than `leaf_size`
//       particles splits into the distinct next-level digits found by binary search in its key
//       range); nodes are stored breadth-first, children of a node contiguous; centre of mass
//       bottom-up in double precision, in a fixed order
//   The resulting arrays are bit-identical to the CPU statement of the same specification
//   (the test oracle; DESIGN.md "Tree specification") — keys, permutation, node ranges, levels, children and {com, mass}.
//
// Traversal (K5), warp-cooperative: a warp owns 32 consecutive targets in key order and walks the
// tree ONCE for the group with a shared stack: every lane tests one node per step against the
// group's bounding box (opening rule of sequential.rs:490-494 with the group's minimum distance,
// so a node is opened whenever ANY member would open it), children are pushed with a warp scan,
// accepted nodes and the particles of opened leaves are appended to a shared interaction list with
// ballot/popc compaction, and whenever 32 entries are ready every lane evaluates all of them for
// its own target (the pair term of gravity/impls/mod.rs:151-166).  A pair at zero distance
// contributes nothing (sequential.rs:485-487 skips a node at the target's position).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "common.cuh"
#include "ptx.cuh"

#include "bh.cuh"

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
static int g_level_build = 0;  // tuning / test hook: 1 = level-wise build instead of the one-pass build,
                               // 2 = one-pass build at every size (also where build_small would run)
static uint32_t g_small_level = 131072;  // levels up to this many nodes take the node-x-digit path

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
// K5a: target groups.  The targets are walked in key order in groups of at most 32 that never
// straddle a coarse cell boundary: a SEGMENT is a maximal cell (key prefix) holding at most
// `seg_max` targets (found from the keys alone: adjacent keys first differ at digit L[i], and the
// cell they share is counted by scanning L to both sides), and every segment is cut into equal
// chunks of <= 32 consecutive targets.  Without this, 32 consecutive keys that cross e.g. the
// centre of a Plummer sphere have a bounding box spanning the core and open millions of nodes.
constexpr int GROUP_BLOCK = 256;
constexpr int SEG_MAX_LIMIT = 1024;

template <int DIM>
__global__ void __launch_bounds__(256) boundary_levels(const uint64_t *__restrict__ keys, int n,
                                                       uint8_t *__restrict__ L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {
        L[0] = 0;
        return;
    }
    const uint64_t x = keys[i] ^ keys[i - 1];
    L[i] = x == 0 ? (uint8_t)(Dims<DIM>::BITS + 1)
                  : (uint8_t)(Dims<DIM>::BITS - (63 - __clzll((long long)x)) / DIM);
}

// Boundary i (between targets i-1 and i) is HARD when the smallest cell holding both targets has
// more than T targets.  That cell spans from the nearest j < i with L[j] < L[i] to the nearest
// k > i with L[k] < L[i] ("nearest smaller value" on both sides; L is 0 outside the array), so
// hard <=> k - j > T.  The walk over L skips 16 entries at a time through a table of chunk
// minima.  Output: one bit per boundary (bits at and past n are set).
constexpr int HARD_CHUNK = 16;
constexpr int HARD_HALO = SEG_MAX_LIMIT + 2 * HARD_CHUNK;

__global__ void __launch_bounds__(GROUP_BLOCK) hard_flags(const uint8_t *__restrict__ L, int n,
                                                          int bits, int T,
                                                          uint32_t *__restrict__ hard_bits) {
    __shared__ __align__(16) uint8_t sL[GROUP_BLOCK + 2 * HARD_HALO];
    __shared__ uint8_t sM[(GROUP_BLOCK + 2 * HARD_HALO) / HARD_CHUNK];
    const int base = blockIdx.x * GROUP_BLOCK;
    const int l0 = base - HARD_HALO;  // global index of sL[0]; a multiple of HARD_CHUNK
    constexpr int NL = GROUP_BLOCK + 2 * HARD_HALO;
    for (int k = threadIdx.x; k < NL; k += GROUP_BLOCK) {
        const int g = l0 + k;
        sL[k] = (g <= 0 || g >= n) ? 0 : L[g];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NL / HARD_CHUNK; c += GROUP_BLOCK) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sL + c * HARD_CHUNK);
        uint32_t m = __vminu4(__vminu4(v.x, v.y), __vminu4(v.z, v.w));
        m = __vminu4(m, m >> 16);
        m = __vminu4(m, m >> 8);
        sM[c] = (uint8_t)(m & 0xffu);
    }
    __syncthreads();
    const int i = base + threadIdx.x;
    bool hard;
    if (i <= 0 || i >= n) hard = true;
    else {
        const int li = sL[i - l0];
        if (li == bits + 1) hard = false;  // identical keys never separate
        else {
            int j = i - 1;  // nearest j < i with L[j] < li (first target of the shared cell)
            const int jmin = i - T - 1;
            while (j > jmin) {
                const int q = j - l0;
                if ((q & (HARD_CHUNK - 1)) == HARD_CHUNK - 1 && sM[q / HARD_CHUNK] >= li) {
                    j -= HARD_CHUNK;
                    continue;
                }
                if (sL[q] < li) break;
                --j;
            }
            if (i - j > T) hard = true;
            else {
                int k = i + 1;  // nearest k > i with L[k] < li (first target past the cell)
                const int kmax = j + T + 1;
                while (k < kmax) {
                    const int q = k - l0;
                    if ((q & (HARD_CHUNK - 1)) == 0 && sM[q / HARD_CHUNK] >= li) {
                        k += HARD_CHUNK;
                        continue;
                    }
                    if (sL[q] < li) break;
                    ++k;
                }
                hard = k - j > T;
            }
        }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, hard);
    if ((threadIdx.x & 31) == 0) hard_bits[i >> 5] = word;
}

// Group starts from the hard boundaries: a SEGMENT runs from one hard boundary to the next; a
// segment of <= T targets is cut into full groups of 32 from its start (the remainder forms one
// small group whose lanes are shared out over the interaction list, see traverse_kernel); longer
// segments (runs of identical keys) are cut at multiples of 32.
__device__ __forceinline__ uint32_t hard_word(const uint32_t *__restrict__ hb, int w, int nwords) {
    return (w < 0 || w >= nwords) ? 0xffffffffu : __ldg(hb + w);
}

__global__ void __launch_bounds__(GROUP_BLOCK) group_flags(const uint32_t *__restrict__ hard_bits,
                                                           int n, int T, int gsize,
                                                           uint32_t *__restrict__ flag) {
    const int i = blockIdx.x * GROUP_BLOCK + threadIdx.x;
    if (i >= n) return;
    const int nwords = (n + 31) >> 5;
    // segment start: nearest hard boundary in [i - T, i]
    int ss = -1;
    {
        int w = i >> 5;
        uint32_t m = hard_word(hard_bits, w, nwords) & (0xffffffffu >> (31 - (i & 31)));
        const int lim = max(i - T, 0);
        for (;;) {
            if (m) {
                const int p = w * 32 + 31 - __clz((int)m);
                if (p >= lim) ss = p;
                break;
            }
            --w;
            if (w * 32 + 31 < lim) break;
            m = hard_word(hard_bits, w, nwords);
        }
    }
    bool start;
    const int gm = gsize - 1;  // gsize = targets per group: 32 or 64
    if (ss < 0) start = (i & gm) == 0;  // inside a long run of identical keys
    else {
        // segment end: next hard boundary in (i, ss + T]
        int se = -1;
        int w = i >> 5;
        uint32_t m = (i & 31) == 31 ? 0u : hard_word(hard_bits, w, nwords) & (0xffffffffu << ((i & 31) + 1));
        const int lim = ss + T;
        for (;;) {
            if (m) {
                const int p = w * 32 + __ffs((int)m) - 1;
                if (p <= lim) se = p;
                break;
            }
            ++w;
            if (w * 32 > lim) break;
            m = hard_word(hard_bits, w, nwords);
        }
        if (se >= 0) start = ((i - ss) & gm) == 0;
        else start = i == ss || (i & gm) == 0;  // the head of a long run of identical keys
    }
    flag[i] = start ? 1u : 0u;
}

// Compaction of the group starts; the last thread also writes the sentinel and the group count.
__global__ void __launch_bounds__(256) scatter_groups(const uint32_t *__restrict__ flag,
                                                      const uint32_t *__restrict__ pos, int n,
                                                      uint32_t *__restrict__ group_start,
                                                      uint32_t *__restrict__ n_groups) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) group_start[pos[i]] = (uint32_t)i;
    if (i == n - 1) {
        const uint32_t g = pos[i] + flag[i];
        group_start[g] = (uint32_t)n;
        *n_groups = g;
    }
}

// ------------------------------------------------------------------------------------------------
// K5b: warp-cooperative theta-traversal (persistent warps, groups handed out by an atomic counter).
constexpr int TRAV_WARPS = 8;      // warps per block
constexpr int STACK_CAP = 1024;    // node indices per warp (shared memory)
constexpr int STACK_RESERVE = 8 * 32;  // room a depth-first descent may still need (7 per level)
constexpr int LIST_CAP = 64;       // interaction ring per warp (float4 entries)
constexpr int MAX_PARTS = 16;      // trees in a forest (= GPUs of a partitioned build)
constexpr int MAX_ROOTS = 736;     // start nodes of a forest walk (STACK_CAP - STACK_RESERVE - 32)

struct TravArgs {
    const NodeRec *nodes;
    const float4 *src;    // sorted sources {x,y,z,mu}
    const float4 *tgt;    // targets in traversal order {x,y,z,_}
    const uint32_t *tgt_perm;  // traversal order -> output row (nullptr: identity)
    const uint32_t *group_start;  // n_groups + 1 entries
    const uint32_t *n_groups;
    uint32_t *work;       // next group to hand out
    float *out;
    unsigned long long *counters;  // [0] node interactions, [1] particle interactions, [2] node tests,
                                   // [3] list entries appended (nodes + particles, per group)
    const Frame *frame;   // root cube extent + mass bound
    int n_tgt;
    int dim;
    float theta2;
    float eps2;
    // Nodes the walk starts from.  nullptr: node 0 (one tree).  Partitioned build (one tree per GPU
    // over the same root cube, see sharded_forest_dev): the root of the merged top tree followed by
    // the loose leaves (partial cells that are leaves in their own part), at most MAX_ROOTS.
    const uint32_t *roots;
    uint32_t n_roots;
};

// The interaction list of a warp lives in shared memory as PAIRS of entries laid out
// {x0 x1 y0 y1}{z0 z1 m0 m1}, so that one lane evaluates two entries at a time with packed FP32
// (FADD2 / FFMA2 / FMUL2): 12 packed + 2 MUFU + 2 LDS.128 per two interactions.
__device__ __forceinline__ void list_store(float *list, int i, const float4 e) {
    float *q = list + (i >> 1) * 8 + (i & 1);
    q[0] = e.x;
    q[2] = e.y;
    q[4] = e.z;
    q[6] = e.w;
}

__device__ __forceinline__ void eval_pair(const float4 A, const float4 B, float2 npx, float2 npy,
                                          float2 npz, float2 eps2p, float2 &ax, float2 &ay,
                                          float2 &az) {
    const float2 dx = ptx::add2(make_float2(A.x, A.y), npx);
    const float2 dy = ptx::add2(make_float2(A.z, A.w), npy);
    const float2 dz = ptx::add2(make_float2(B.x, B.y), npz);
    float2 r2 = ptx::fma2(dx, dx, eps2p);
    r2 = ptx::fma2(dy, dy, r2);
    r2 = ptx::fma2(dz, dz, r2);
    // zero distance contributes nothing: eps2p carries, on top of the softening, a floor t chosen
    // so that (largest node mass) * r^-3 stays finite, hence the term is d * finite = 0; t is far
    // below the resolution of distinct f32 positions (r2 + t == r2 bit for bit for r2 >= 2^24 t)
    float2 ri;
    ri.x = ptx::rsqrt_approx(r2.x);
    ri.y = ptx::rsqrt_approx(r2.y);
    const float2 ri2 = ptx::mul2(ri, ri);
    const float2 mri = ptx::mul2(ri, make_float2(B.z, B.w));
    const float2 sc = ptx::mul2(ri2, mri);
    ax = ptx::fma2(dx, sc, ax);
    ay = ptx::fma2(dy, sc, ay);
    az = ptx::fma2(dz, sc, az);
}

template <bool COUNT>
__global__ void __launch_bounds__(TRAV_WARPS * 32, 4) traverse_kernel(TravArgs a) {
    __shared__ uint32_t s_stack[TRAV_WARPS][STACK_CAP];
    __shared__ __align__(16) float4 s_list[TRAV_WARPS][LIST_CAP];

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stack = s_stack[warp];
    float4 *list4 = s_list[warp];
    float *list = reinterpret_cast<float *>(list4);
    const uint32_t n_groups = *a.n_groups;
    const float ext = a.frame->ext;
    // r2 floor such that (largest node mass) * r^-3 stays finite (see eval_pair)
    const float cb = cbrtf(fminf(a.frame->mass_bound, 3e38f)) * 2.2e-13f;
    const float tiny = fmaxf(2.f * cb * cb, 1e-36f);
    const float2 eps2p = make_float2(a.eps2 + tiny, a.eps2 + tiny);
    unsigned long long c_node = 0, c_part = 0, c_test = 0, c_entries = 0;

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const int t0 = (int)a.group_start[g];
        const int gcnt = (int)a.group_start[g + 1] - t0;  // 1..32 targets
        // lanes = (target, slice): a group of <= 16 targets uses 32 / gpad lanes per target, each
        // evaluating every (32 / gpad)-th interaction; partial sums are combined at the end
        int gpad = 1;
        while (gpad < gcnt) gpad <<= 1;
        const int slices = 32 / gpad;
        const int tl = lane & (gpad - 1), slice = lane / gpad;
        const int ti = t0 + min(tl, gcnt - 1);
        const float4 tp = a.tgt[ti];
        const float px = tp.x, py = tp.y, pz = tp.z;

        // group bounding box -> centre and half extent
        float lox = px, hix = px, loy = py, hiy = py, loz = pz, hiz = pz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lox = fminf(lox, __shfl_xor_sync(FULL, lox, o));
            hix = fmaxf(hix, __shfl_xor_sync(FULL, hix, o));
            loy = fminf(loy, __shfl_xor_sync(FULL, loy, o));
            hiy = fmaxf(hiy, __shfl_xor_sync(FULL, hiy, o));
            loz = fminf(loz, __shfl_xor_sync(FULL, loz, o));
            hiz = fmaxf(hiz, __shfl_xor_sync(FULL, hiz, o));
        }
        const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
        const float hx = 0.5f * (hix - lox), hy = 0.5f * (hiy - loy), hz = 0.5f * (hiz - loz);

        const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py),
                     npz = make_float2(-pz, -pz);
        float2 ax2 = make_float2(0.f, 0.f), ay2 = ax2, az2 = ax2;
        unsigned long long g_node = 0, g_part = 0;
        int sp = 1;    // stack size (uniform across the warp)
        int fill = 0;  // entries in the interaction list (uniform)
        __syncwarp();
        if (lane == 0) stack[0] = 0;
        __syncwarp();

        auto flush_full = [&]() {  // evaluate the first 32 entries once they are ready
            if (fill >= 32) {
                __syncwarp();
                if (slices == 1) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        eval_pair(list4[2 * q], list4[2 * q + 1], npx, npy, npz, eps2p, ax2, ay2, az2);
                } else {
                    for (int q = slice; q < 16; q += slices)
                        eval_pair(list4[2 * q], list4[2 * q + 1], npx, npy, npz, eps2p, ax2, ay2, az2);
                }
                fill -= 32;
                // move the remainder (< 32 entries = <= 16 pairs = <= 32 float4) to the front
                const bool mv = lane < ((fill + 1) >> 1) * 2;
                float4 v;
                if (mv) v = list4[32 + lane];
                __syncwarp();
                if (mv) list4[lane] = v;
                __syncwarp();
            }
        };

        while (sp > 0) {
            // pop up to 32 nodes, but never so many that their children could overflow the stack
            const int room = (STACK_CAP - STACK_RESERVE - sp) / 7;
            const int k = min(min(32, sp), max(room, 1));
            const bool has = lane < k;
            NodeRec nd;
            nd.cm = make_float4(0.f, 0.f, 0.f, 0.f);
            nd.first_child = 0;
            nd.begin = 0;
            nd.count = 0;
            nd.nchild_level = 0;
            if (has) {
                const uint32_t id = stack[sp - 1 - lane];
                const uint4 *q = reinterpret_cast<const uint4 *>(a.nodes + id);
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
                nd.cm = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y),
                                    __uint_as_float(q0.z), __uint_as_float(q0.w));
                nd.first_child = q1.x;
                nd.nchild_level = q1.y;
                nd.begin = q1.z;
                nd.count = q1.w;
            }
            sp -= k;
            __syncwarp();

            // opening rule for the group: (theta^2) * dmin^2 < width^2, dmin = distance from the
            // centre of mass to the group's bounding box
            bool open = false;
            if (has) {
                const float ddx = fmaxf(fabsf(nd.cm.x - cx) - hx, 0.f);
                const float ddy = fmaxf(fabsf(nd.cm.y - cy) - hy, 0.f);
                const float ddz = fmaxf(fabsf(nd.cm.z - cz) - hz, 0.f);
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                const int level = (int)(nd.nchild_level >> 8);
                const float w = ext * __int_as_float((127 - level) << 23);
                open = a.theta2 * d2 < w * w;
            }
            const uint32_t nc = nd.nchild_level & 0xffu;
            const bool open_internal = has && open && nc > 0;
            const bool open_leaf = has && open && nc == 0;
            const bool accept = has && !open && nd.cm.w != 0.f;
            if (COUNT) c_test += k;

            // one warp scan serves both the children to push (low 10 bits, <= 256 in total) and
            // the particles of opened leaves (high 22 bits); a leaf too large for the packing
            // (only possible at the last level, many identical keys) takes a second scan
            const int c_child = open_internal ? (int)nc : 0;
            const int c_leaf = open_leaf ? (int)nd.count : 0;
            const bool wide = __any_sync(FULL, c_leaf > 65535);
            int leaf_incl;
            {
                unsigned packed = (unsigned)c_child | (wide ? 0u : (unsigned)c_leaf << 10);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(FULL, packed, o);
                    if (lane >= o) packed += v;
                }
                const int incl = (int)(packed & 1023u);
                leaf_incl = (int)(packed >> 10);
                const int total = __shfl_sync(FULL, incl, 31);
                const int base = sp + incl - c_child;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < c_child) stack[base + j] = nd.first_child + j;
                sp += total;
            }
            if (wide) {
                leaf_incl = c_leaf;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, leaf_incl, o);
                    if (lane >= o) leaf_incl += v;
                }
            }

            // accepted nodes -> interaction ring
            {
                const unsigned m = __ballot_sync(FULL, accept);
                if (m) {
                    if (accept) list_store(list, fill + __popc(m & ((1u << lane) - 1)), nd.cm);
                    const int cnt = __popc(m);
                    if (COUNT) g_node += cnt;
                    fill += cnt;
                    flush_full();
                }
            }

            // particles of opened leaves -> interaction ring, 32 particles per round: lane f of a
            // round finds the leaf that owns flat index f by a shuffle binary search over the
            // inclusive scan of the leaf sizes, so every round is one coalesced-per-leaf load
            {
                const int incl = leaf_incl;
                const int total = __shfl_sync(FULL, incl, 31);
                const int excl = incl - c_leaf;
                for (int base = 0; base < total; base += 32) {
                    const int f = base + lane;
                    int owner = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(FULL, incl, (owner + step - 1) & 31);
                        if (v <= f) owner += step;
                    }
                    owner = min(owner, 31);
                    const uint32_t ob = __shfl_sync(FULL, nd.begin, owner);
                    const int oe = __shfl_sync(FULL, excl, owner);
                    if (f < total) list_store(list, fill + lane, __ldg(a.src + ob + (f - oe)));
                    const int cnt = min(32, total - base);
                    if (COUNT) g_part += cnt;
                    fill += cnt;
                    flush_full();
                }
            }
            __syncwarp();
        }
        if (fill > 0) {
            if ((fill & 1) && lane == 0) list_store(list, fill, make_float4(0.f, 0.f, 0.f, 0.f));
            __syncwarp();
            const int pairs = (fill + 1) >> 1;
            for (int q = slice; q < pairs; q += slices)
                eval_pair(list4[2 * q], list4[2 * q + 1], npx, npy, npz, eps2p, ax2, ay2, az2);
        }
        float ax = ax2.x + ax2.y, ay = ay2.x + ay2.y, az = az2.x + az2.y;
        for (int o = gpad; o < 32; o <<= 1) {  // combine the slices of each target
            ax += __shfl_xor_sync(FULL, ax, o);
            ay += __shfl_xor_sync(FULL, ay, o);
            az += __shfl_xor_sync(FULL, az, o);
        }
        if (slice == 0 && tl < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ti] : (uint32_t)ti;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = ax;
            o[1] = ay;
            if (a.dim == 3) o[2] = az;
        }
        if (COUNT) {  // per-target counts: every target of the group saw every list entry
            c_node += g_node * gcnt;
            c_part += g_part * gcnt;
            c_entries += g_node + g_part;
        }
    }
    if (COUNT && lane == 0) {
        atomicAdd(a.counters + 0, c_node);
        atomicAdd(a.counters + 1, c_part);
        atomicAdd(a.counters + 2, c_test);
        atomicAdd(a.counters + 5, c_entries);
    }
}

// ------------------------------------------------------------------------------------------------
// K5c: the same walk with TWO targets per lane (groups of up to 64 targets).  The packed FP32
// lanes now hold two targets and an interaction-list entry is a scalar-broadcast operand
// (FADD2 Rd, -Rtargets.F32x2, Rentry.F32), so the list is plain {x,y,z,mu} records: one
// conflict-free STS.128 per appended entry, one broadcast LDS.128 per entry and pair of targets,
// and the tree walk is shared by twice as many targets.  The list is a 64-entry ring that is
// evaluated 32 entries at a time.
__device__ __forceinline__ void eval_entry(const float4 e, float2 npx, float2 npy, float2 npz,
                                           float2 eps2p, float2 &ax, float2 &ay, float2 &az) {
    const float2 dx = ptx::add2(ptx::splat(e.x), npx);
    const float2 dy = ptx::add2(ptx::splat(e.y), npy);
    const float2 dz = ptx::add2(ptx::splat(e.z), npz);
    float2 r2 = ptx::fma2(dx, dx, eps2p);
    r2 = ptx::fma2(dy, dy, r2);
    r2 = ptx::fma2(dz, dz, r2);
    float2 ri;
    ri.x = ptx::rsqrt_approx(r2.x);
    ri.y = ptx::rsqrt_approx(r2.y);
    const float2 ri2 = ptx::mul2(ri, ri);
    const float2 mri = ptx::mul2(ri, ptx::splat(e.w));
    const float2 sc = ptx::mul2(ri2, mri);
    ax = ptx::fma2(dx, sc, ax);
    ay = ptx::fma2(dy, sc, ay);
    az = ptx::fma2(dz, sc, az);
}

// VAR: experiment bits (tuning only).  1 = walk only (no evaluation), 2 = prefetch the next round's
// node records into L1 before the evaluation.
// FOREST: the walk starts from a.roots[0 .. a.n_roots) (partitioned multi-GPU build) instead of node 0.
template <bool COUNT, int VAR = 0, bool FOREST = false>
__global__ void __launch_bounds__(TRAV_WARPS * 32, 4) traverse2_kernel(TravArgs a) {
    __shared__ uint32_t s_stack[TRAV_WARPS][STACK_CAP];
    __shared__ __align__(16) float4 s_list[TRAV_WARPS][LIST_CAP];

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stack = s_stack[warp];
    float4 *list4 = s_list[warp];
    const uint32_t n_groups = *a.n_groups;
    const float ext = a.frame->ext;
    const float cb = cbrtf(fminf(a.frame->mass_bound, 3e38f)) * 2.2e-13f;
    const float tiny = fmaxf(2.f * cb * cb, 1e-36f);  // see eval_pair
    const float2 eps2p = make_float2(a.eps2 + tiny, a.eps2 + tiny);
    unsigned long long c_node = 0, c_part = 0, c_test = 0, c_entries = 0;

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const int t0 = (int)a.group_start[g];
        const int gcnt = (int)a.group_start[g + 1] - t0;  // 1..64 targets
        // lanes = (pair of targets, slice): a group of <= 32 targets uses 64 / gpad lanes per
        // pair, each evaluating every (64 / gpad)-th list entry; partial sums are combined at
        // the end
        int half = 1;  // lanes per slice = gpad / 2
        while (2 * half < gcnt) half <<= 1;
        const int slices = 32 / half;
        const int tl = lane & (half - 1), slice = lane / half;
        const int ia = t0 + min(tl, gcnt - 1), ib = t0 + min(tl + half, gcnt - 1);
        // scalar loads on purpose: each (a, b) coordinate pair is then free to land in an aligned
        // register pair, the operand form of FADD2; out of two LDG.128 quads ptxas re-packs the
        // pair with two MOVs in front of every FADD2
        float3 ta, tb;
        ta.x = ptx::ldg_f32(&a.tgt[ia].x);
        tb.x = ptx::ldg_f32(&a.tgt[ib].x);
        ta.y = ptx::ldg_f32(&a.tgt[ia].y);
        tb.y = ptx::ldg_f32(&a.tgt[ib].y);
        ta.z = ptx::ldg_f32(&a.tgt[ia].z);
        tb.z = ptx::ldg_f32(&a.tgt[ib].z);

        float lox = fminf(ta.x, tb.x), hix = fmaxf(ta.x, tb.x);
        float loy = fminf(ta.y, tb.y), hiy = fmaxf(ta.y, tb.y);
        float loz = fminf(ta.z, tb.z), hiz = fmaxf(ta.z, tb.z);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lox = fminf(lox, __shfl_xor_sync(FULL, lox, o));
            hix = fmaxf(hix, __shfl_xor_sync(FULL, hix, o));
            loy = fminf(loy, __shfl_xor_sync(FULL, loy, o));
            hiy = fmaxf(hiy, __shfl_xor_sync(FULL, hiy, o));
            loz = fminf(loz, __shfl_xor_sync(FULL, loz, o));
            hiz = fmaxf(hiz, __shfl_xor_sync(FULL, hiz, o));
        }
        const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
        const float hx = 0.5f * (hix - lox), hy = 0.5f * (hiy - loy), hz = 0.5f * (hiz - loz);

        const float2 npx = make_float2(-ta.x, -tb.x), npy = make_float2(-ta.y, -tb.y),
                     npz = make_float2(-ta.z, -tb.z);
        float2 ax2 = make_float2(0.f, 0.f), ay2 = ax2, az2 = ax2;
        unsigned long long g_node = 0, g_part = 0;
        int sp = 1;    // stack size (uniform across the warp)
        int head = 0;  // ring position of the oldest list entry: 0 or 32 (uniform)
        int fill = 0;  // entries in the ring (uniform), < 32 between steps
        __syncwarp();
        if (FOREST) {
            sp = (int)a.n_roots;
            for (int i = lane; i < sp; i += 32) stack[i] = a.roots[i];
        } else if (lane == 0) {
            stack[0] = 0;
        }
        __syncwarp();

        auto flush_full = [&]() {  // evaluate the 32 oldest entries once they are ready
            if (fill >= 32) {
                __syncwarp();
                const float4 *blk = list4 + head;
                if (VAR & 1) {
                } else if (slices == 1) {
#pragma unroll 16
                    for (int q = 0; q < 32; ++q) eval_entry(blk[q], npx, npy, npz, eps2p, ax2, ay2, az2);
                } else {
                    for (int q = slice; q < 32; q += slices)
                        eval_entry(blk[q], npx, npy, npz, eps2p, ax2, ay2, az2);
                }
                fill -= 32;
                head ^= 32;
                __syncwarp();
            }
        };

        while (sp > 0) {
            const int room = (STACK_CAP - STACK_RESERVE - sp) / 7;
            const int k = min(min(32, sp), max(room, 1));
            const bool has = lane < k;
            NodeRec nd;
            nd.cm = make_float4(0.f, 0.f, 0.f, 0.f);
            nd.first_child = 0;
            nd.begin = 0;
            nd.count = 0;
            nd.nchild_level = 0;
            if (has) {
                const uint32_t id = stack[sp - 1 - lane];
                const uint4 *q = reinterpret_cast<const uint4 *>(a.nodes + id);
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
                nd.cm = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y),
                                    __uint_as_float(q0.z), __uint_as_float(q0.w));
                nd.first_child = q1.x;
                nd.nchild_level = q1.y;
                nd.begin = q1.z;
                nd.count = q1.w;
            }
            sp -= k;
            __syncwarp();

            bool open = false;
            if (has) {
                const float ddx = fmaxf(fabsf(nd.cm.x - cx) - hx, 0.f);
                const float ddy = fmaxf(fabsf(nd.cm.y - cy) - hy, 0.f);
                const float ddz = fmaxf(fabsf(nd.cm.z - cz) - hz, 0.f);
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                const int level = (int)(nd.nchild_level >> 8);
                const float w = ext * __int_as_float((127 - level) << 23);
                open = a.theta2 * d2 < w * w;
            }
            const uint32_t nc = nd.nchild_level & 0xffu;
            const bool open_internal = has && open && nc > 0;
            const bool open_leaf = has && open && nc == 0;
            const bool accept = has && !open && nd.cm.w != 0.f;
            if (COUNT) c_test += k;

            const int c_child = open_internal ? (int)nc : 0;
            const int c_leaf = open_leaf ? (int)nd.count : 0;
            const bool wide = __any_sync(FULL, c_leaf > 65535);
            int leaf_incl;
            {
                unsigned packed = (unsigned)c_child | (wide ? 0u : (unsigned)c_leaf << 10);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned v = __shfl_up_sync(FULL, packed, o);
                    if (lane >= o) packed += v;
                }
                const int incl = (int)(packed & 1023u);
                leaf_incl = (int)(packed >> 10);
                const int total = __shfl_sync(FULL, incl, 31);
                const int base = sp + incl - c_child;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < c_child) stack[base + j] = nd.first_child + j;
                sp += total;
            }
            if (VAR & 2) {  // the next round's nodes are known now: pull them into L1
                __syncwarp();
                if (lane < sp)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nodes + stack[sp - 1 - lane]));
            }
            if (wide) {
                leaf_incl = c_leaf;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, leaf_incl, o);
                    if (lane >= o) leaf_incl += v;
                }
            }

            {  // accepted nodes -> ring
                const unsigned m = __ballot_sync(FULL, accept);
                if (m) {
                    if (accept)
                        list4[(head + fill + __popc(m & ((1u << lane) - 1))) & (LIST_CAP - 1)] = nd.cm;
                    const int cnt = __popc(m);
                    if (COUNT) g_node += cnt;
                    fill += cnt;
                    flush_full();
                }
            }

            {  // particles of opened leaves -> ring, 32 per round (see traverse_kernel)
                const int incl = leaf_incl;
                const int total = __shfl_sync(FULL, incl, 31);
                const int excl = incl - c_leaf;
                for (int base = 0; base < total; base += 32) {
                    const int f = base + lane;
                    int owner = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(FULL, incl, (owner + step - 1) & 31);
                        if (v <= f) owner += step;
                    }
                    owner = min(owner, 31);
                    const uint32_t ob = __shfl_sync(FULL, nd.begin, owner);
                    const int oe = __shfl_sync(FULL, excl, owner);
                    if (f < total)
                        list4[(head + fill + lane) & (LIST_CAP - 1)] = __ldg(a.src + ob + (f - oe));
                    const int cnt = min(32, total - base);
                    if (COUNT) g_part += cnt;
                    fill += cnt;
                    flush_full();
                }
            }
            __syncwarp();
        }
        if (fill > 0 && !(VAR & 1)) {
            __syncwarp();
            for (int q = slice; q < fill; q += slices)
                eval_entry(list4[(head + q) & (LIST_CAP - 1)], npx, npy, npz, eps2p, ax2, ay2, az2);
        }
        float axa = ax2.x, aya = ay2.x, aza = az2.x, axb = ax2.y, ayb = ay2.y, azb = az2.y;
        for (int o = half; o < 32; o <<= 1) {  // combine the slices of each target
            axa += __shfl_xor_sync(FULL, axa, o);
            aya += __shfl_xor_sync(FULL, aya, o);
            aza += __shfl_xor_sync(FULL, aza, o);
            axb += __shfl_xor_sync(FULL, axb, o);
            ayb += __shfl_xor_sync(FULL, ayb, o);
            azb += __shfl_xor_sync(FULL, azb, o);
        }
        if (slice == 0 && tl < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ia] : (uint32_t)ia;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = axa;
            o[1] = aya;
            if (a.dim == 3) o[2] = aza;
        }
        if (slice == 0 && tl + half < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ib] : (uint32_t)ib;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = axb;
            o[1] = ayb;
            if (a.dim == 3) o[2] = azb;
        }
        if (COUNT) {
            c_node += g_node * gcnt;
            c_part += g_part * gcnt;
            c_entries += g_node + g_part;
        }
    }
    if (COUNT && lane == 0) {
        atomicAdd(a.counters + 0, c_node);
        atomicAdd(a.counters + 1, c_part);
        atomicAdd(a.counters + 2, c_test);
        atomicAdd(a.counters + 5, c_entries);
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

// monopole + quadrupole term of one node for the two targets of a lane (packed FP32)
__device__ __forceinline__ void eval_node_q(const float4 c, const float4 qa, const float4 qb,
                                            float2 npx, float2 npy, float2 npz, float2 eps2p,
                                            float2 &ax, float2 &ay, float2 &az) {
    const float2 dx = ptx::add2(ptx::splat(c.x), npx);
    const float2 dy = ptx::add2(ptx::splat(c.y), npy);
    const float2 dz = ptx::add2(ptx::splat(c.z), npz);
    float2 r2 = ptx::fma2(dx, dx, eps2p);
    r2 = ptx::fma2(dy, dy, r2);
    r2 = ptx::fma2(dz, dz, r2);
    float2 ri;
    ri.x = ptx::rsqrt_approx(r2.x);
    ri.y = ptx::rsqrt_approx(r2.y);
    const float2 ri2 = ptx::mul2(ri, ri);
    const float2 ux = ptx::mul2(dx, ri), uy = ptx::mul2(dy, ri), uz = ptx::mul2(dz, ri);
    // Qu' = (Q u) ri^2
    float2 qx = ptx::mul2(ptx::splat(qa.x), ux);
    qx = ptx::fma2(ptx::splat(qa.y), uy, qx);
    qx = ptx::fma2(ptx::splat(qa.z), uz, qx);
    float2 qy = ptx::mul2(ptx::splat(qa.y), ux);
    qy = ptx::fma2(ptx::splat(qa.w), uy, qy);
    qy = ptx::fma2(ptx::splat(qb.x), uz, qy);
    float2 qz = ptx::mul2(ptx::splat(qa.z), ux);
    qz = ptx::fma2(ptx::splat(qb.x), uy, qz);
    qz = ptx::fma2(ptx::splat(qb.y), uz, qz);
    qx = ptx::mul2(qx, ri2);
    qy = ptx::mul2(qy, ri2);
    qz = ptx::mul2(qz, ri2);
    float2 uqu = ptx::mul2(ux, qx);
    uqu = ptx::fma2(uy, qy, uqu);
    uqu = ptx::fma2(uz, qz, uqu);
    // s = M + 5/2 u.Qu'   (coefficient of u; everything is multiplied by ri^2 at the end)
    const float2 s = ptx::fma2(ptx::splat(2.5f), uqu, ptx::splat(c.w));
    const float2 vx = ptx::fma2(s, ux, ptx::mul2(qx, ptx::splat(-1.f)));
    const float2 vy = ptx::fma2(s, uy, ptx::mul2(qy, ptx::splat(-1.f)));
    const float2 vz = ptx::fma2(s, uz, ptx::mul2(qz, ptx::splat(-1.f)));
    ax = ptx::fma2(vx, ri2, ax);
    ay = ptx::fma2(vy, ri2, ay);
    az = ptx::fma2(vz, ri2, az);
}

constexpr int TRAVQ_WARPS = 4;

// traverse2_kernel with quadrupole nodes: accepted nodes go to their own ring ({com, mass} + two
// quadrupole quads per entry), the particles of opened leaves to the plain ring; either ring is
// evaluated 32 entries at a time.  A node whose centre of mass touches the group's box is opened
// whatever theta says (the expansion is singular at zero distance).
__global__ void __launch_bounds__(TRAVQ_WARPS * 32) traverse2q_kernel(TravArgs a, const float4 *__restrict__ quad) {
    __shared__ uint32_t s_stack[TRAVQ_WARPS][STACK_CAP];
    __shared__ __align__(16) float4 s_list[TRAVQ_WARPS][LIST_CAP];
    __shared__ __align__(16) float4 s_nodes[TRAVQ_WARPS][3 * LIST_CAP];

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stack = s_stack[warp];
    float4 *list4 = s_list[warp];
    float4 *nlist = s_nodes[warp];
    const uint32_t n_groups = *a.n_groups;
    const float ext = a.frame->ext;
    const float cb = cbrtf(fminf(a.frame->mass_bound, 3e38f)) * 2.2e-13f;
    const float tiny = fmaxf(2.f * cb * cb, 1e-36f);
    const float2 eps2p = make_float2(a.eps2 + tiny, a.eps2 + tiny);

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const int t0 = (int)a.group_start[g];
        const int gcnt = (int)a.group_start[g + 1] - t0;  // 1..64 targets
        int half = 1;
        while (2 * half < gcnt) half <<= 1;
        const int slices = 32 / half;
        const int tl = lane & (half - 1), slice = lane / half;
        const int ia = t0 + min(tl, gcnt - 1), ib = t0 + min(tl + half, gcnt - 1);
        float3 ta, tb;
        ta.x = ptx::ldg_f32(&a.tgt[ia].x);
        tb.x = ptx::ldg_f32(&a.tgt[ib].x);
        ta.y = ptx::ldg_f32(&a.tgt[ia].y);
        tb.y = ptx::ldg_f32(&a.tgt[ib].y);
        ta.z = ptx::ldg_f32(&a.tgt[ia].z);
        tb.z = ptx::ldg_f32(&a.tgt[ib].z);

        float lox = fminf(ta.x, tb.x), hix = fmaxf(ta.x, tb.x);
        float loy = fminf(ta.y, tb.y), hiy = fmaxf(ta.y, tb.y);
        float loz = fminf(ta.z, tb.z), hiz = fmaxf(ta.z, tb.z);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lox = fminf(lox, __shfl_xor_sync(FULL, lox, o));
            hix = fmaxf(hix, __shfl_xor_sync(FULL, hix, o));
            loy = fminf(loy, __shfl_xor_sync(FULL, loy, o));
            hiy = fmaxf(hiy, __shfl_xor_sync(FULL, hiy, o));
            loz = fminf(loz, __shfl_xor_sync(FULL, loz, o));
            hiz = fmaxf(hiz, __shfl_xor_sync(FULL, hiz, o));
        }
        const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
        const float hx = 0.5f * (hix - lox), hy = 0.5f * (hiy - loy), hz = 0.5f * (hiz - loz);

        const float2 npx = make_float2(-ta.x, -tb.x), npy = make_float2(-ta.y, -tb.y),
                     npz = make_float2(-ta.z, -tb.z);
        float2 ax2 = make_float2(0.f, 0.f), ay2 = ax2, az2 = ax2;
        int sp = 1;
        int head = 0, fill = 0;    // particle ring
        int nhead = 0, nfill = 0;  // node ring
        __syncwarp();
        if (lane == 0) stack[0] = 0;
        __syncwarp();

        auto flush_particles = [&]() {
            if (fill >= 32) {
                __syncwarp();
                const float4 *blk = list4 + head;
                for (int q = slice; q < 32; q += slices) eval_entry(blk[q], npx, npy, npz, eps2p, ax2, ay2, az2);
                fill -= 32;
                head ^= 32;
                __syncwarp();
            }
        };
        auto flush_nodes = [&]() {
            if (nfill >= 32) {
                __syncwarp();
                const float4 *blk = nlist + 3 * nhead;
                for (int q = slice; q < 32; q += slices)
                    eval_node_q(blk[3 * q], blk[3 * q + 1], blk[3 * q + 2], npx, npy, npz, eps2p, ax2, ay2, az2);
                nfill -= 32;
                nhead ^= 32;
                __syncwarp();
            }
        };

        while (sp > 0) {
            const int room = (STACK_CAP - STACK_RESERVE - sp) / 7;
            const int k = min(min(32, sp), max(room, 1));
            const bool has = lane < k;
            NodeRec nd;
            nd.cm = make_float4(0.f, 0.f, 0.f, 0.f);
            nd.first_child = 0;
            nd.begin = 0;
            nd.count = 0;
            nd.nchild_level = 0;
            uint32_t id = 0;
            if (has) {
                id = stack[sp - 1 - lane];
                const uint4 *q = reinterpret_cast<const uint4 *>(a.nodes + id);
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
                nd.cm = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y),
                                    __uint_as_float(q0.z), __uint_as_float(q0.w));
                nd.first_child = q1.x;
                nd.nchild_level = q1.y;
                nd.begin = q1.z;
                nd.count = q1.w;
            }
            sp -= k;
            __syncwarp();

            bool open = false;
            if (has) {
                const float ddx = fmaxf(fabsf(nd.cm.x - cx) - hx, 0.f);
                const float ddy = fmaxf(fabsf(nd.cm.y - cy) - hy, 0.f);
                const float ddz = fmaxf(fabsf(nd.cm.z - cz) - hz, 0.f);
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                const int level = (int)(nd.nchild_level >> 8);
                const float w = ext * __int_as_float((127 - level) << 23);
                open = a.theta2 * d2 < w * w || d2 == 0.f;
            }
            const uint32_t nc = nd.nchild_level & 0xffu;
            const bool open_internal = has && open && nc > 0;
            const bool open_leaf = has && open && nc == 0;
            const bool accept = has && !open && nd.cm.w != 0.f;

            const int c_child = open_internal ? (int)nc : 0;
            const int c_leaf = open_leaf ? (int)nd.count : 0;
            int child_incl = c_child, leaf_incl = c_leaf;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, child_incl, o);
                const int u = __shfl_up_sync(FULL, leaf_incl, o);
                if (lane >= o) {
                    child_incl += v;
                    leaf_incl += u;
                }
            }
            {
                const int total = __shfl_sync(FULL, child_incl, 31);
                const int base = sp + child_incl - c_child;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < c_child) stack[base + j] = nd.first_child + j;
                sp += total;
            }

            {  // accepted nodes -> node ring
                const unsigned m = __ballot_sync(FULL, accept);
                if (m) {
                    if (accept) {
                        const int slot = (nhead + nfill + __popc(m & ((1u << lane) - 1))) & (LIST_CAP - 1);
                        nlist[3 * slot] = nd.cm;
                        nlist[3 * slot + 1] = __ldg(quad + 2 * (size_t)id);
                        nlist[3 * slot + 2] = __ldg(quad + 2 * (size_t)id + 1);
                    }
                    nfill += __popc(m);
                    flush_nodes();
                }
            }

            {  // particles of opened leaves -> particle ring, 32 per round
                const int incl = leaf_incl;
                const int total = __shfl_sync(FULL, incl, 31);
                const int excl = incl - c_leaf;
                for (int base = 0; base < total; base += 32) {
                    const int f = base + lane;
                    int owner = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(FULL, incl, (owner + step - 1) & 31);
                        if (v <= f) owner += step;
                    }
                    owner = min(owner, 31);
                    const uint32_t ob = __shfl_sync(FULL, nd.begin, owner);
                    const int oe = __shfl_sync(FULL, excl, owner);
                    if (f < total)
                        list4[(head + fill + lane) & (LIST_CAP - 1)] = __ldg(a.src + ob + (f - oe));
                    fill += min(32, total - base);
                    flush_particles();
                }
            }
            __syncwarp();
        }
        __syncwarp();
        for (int q = slice; q < fill; q += slices)
            eval_entry(list4[(head + q) & (LIST_CAP - 1)], npx, npy, npz, eps2p, ax2, ay2, az2);
        for (int q = slice; q < nfill; q += slices) {
            const int slot = (nhead + q) & (LIST_CAP - 1);
            eval_node_q(nlist[3 * slot], nlist[3 * slot + 1], nlist[3 * slot + 2], npx, npy, npz, eps2p, ax2,
                        ay2, az2);
        }
        float axa = ax2.x, aya = ay2.x, aza = az2.x, axb = ax2.y, ayb = ay2.y, azb = az2.y;
        for (int o = half; o < 32; o <<= 1) {
            axa += __shfl_xor_sync(FULL, axa, o);
            aya += __shfl_xor_sync(FULL, aya, o);
            aza += __shfl_xor_sync(FULL, aza, o);
            axb += __shfl_xor_sync(FULL, axb, o);
            ayb += __shfl_xor_sync(FULL, ayb, o);
            azb += __shfl_xor_sync(FULL, azb, o);
        }
        if (slice == 0 && tl < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ia] : (uint32_t)ia;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = axa;
            o[1] = aya;
            if (a.dim == 3) o[2] = aza;
        }
        if (slice == 0 && tl + half < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ib] : (uint32_t)ib;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = axb;
            o[1] = ayb;
            if (a.dim == 3) o[2] = azb;
        }
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

struct Ext64 {
    const double4 *src64;  // sources in key order
    const double4 *cm64;   // {com, mass} per node
    const double4 *tgt64;  // targets in traversal order {x, y, z|0, _}
    double *out;
    double eps2;
};

constexpr int TRAV64_WARPS = 4;

__device__ __forceinline__ void eval_entry64(const double4 e, double px, double py, double pz,
                                             double eps2, double &ax, double &ay, double &az) {
    const double dx = e.x - px, dy = e.y - py, dz = e.z - pz;
    double r2 = fma(dx, dx, eps2);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    r2 = ptx::one_if_zero(r2);  // zero distance: d == 0, so the term is 0 * finite = 0
    const double sc = ptx::mu_rcbrt2(r2, e.w);
    ax = fma(dx, sc, ax);
    ay = fma(dy, sc, ay);
    az = fma(dz, sc, az);
}

// The walk of traverse2_kernel (shared stack, group bounding box, ring of list entries), one
// target per lane, groups of <= 32, entries and arithmetic in double precision.
__global__ void __launch_bounds__(TRAV64_WARPS * 32) traverse64_kernel(TravArgs a, Ext64 x) {
    __shared__ uint32_t s_stack[TRAV64_WARPS][STACK_CAP];
    __shared__ __align__(16) double4 s_list[TRAV64_WARPS][LIST_CAP];

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stack = s_stack[warp];
    double4 *list = s_list[warp];
    const uint32_t n_groups = *a.n_groups;
    const float ext = a.frame->ext;

    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.work, 1u);
        g = __shfl_sync(FULL, g, 0);
        if (g >= n_groups) break;
        const int t0 = (int)a.group_start[g];
        const int gcnt = (int)a.group_start[g + 1] - t0;  // 1..32 targets
        int gpad = 1;
        while (gpad < gcnt) gpad <<= 1;
        const int slices = 32 / gpad;
        const int tl = lane & (gpad - 1), slice = lane / gpad;
        const int ti = t0 + min(tl, gcnt - 1);
        const double4 tp = x.tgt64[ti];
        const double px = tp.x, py = tp.y, pz = tp.z;

        // group bounding box in f32, rounded outwards
        float lox = __double2float_rd(px), hix = __double2float_ru(px);
        float loy = __double2float_rd(py), hiy = __double2float_ru(py);
        float loz = __double2float_rd(pz), hiz = __double2float_ru(pz);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lox = fminf(lox, __shfl_xor_sync(FULL, lox, o));
            hix = fmaxf(hix, __shfl_xor_sync(FULL, hix, o));
            loy = fminf(loy, __shfl_xor_sync(FULL, loy, o));
            hiy = fmaxf(hiy, __shfl_xor_sync(FULL, hiy, o));
            loz = fminf(loz, __shfl_xor_sync(FULL, loz, o));
            hiz = fmaxf(hiz, __shfl_xor_sync(FULL, hiz, o));
        }
        const float cx = 0.5f * (lox + hix), cy = 0.5f * (loy + hiy), cz = 0.5f * (loz + hiz);
        const float hx = 0.5f * (hix - lox), hy = 0.5f * (hiy - loy), hz = 0.5f * (hiz - loz);

        double ax = 0.0, ay = 0.0, az = 0.0;
        int sp = 1, head = 0, fill = 0;
        __syncwarp();
        if (lane == 0) stack[0] = 0;
        __syncwarp();

        auto flush_full = [&]() {
            if (fill >= 32) {
                __syncwarp();
                const double4 *blk = list + head;
                for (int q = slice; q < 32; q += slices) eval_entry64(blk[q], px, py, pz, x.eps2, ax, ay, az);
                fill -= 32;
                head ^= 32;
                __syncwarp();
            }
        };

        while (sp > 0) {
            const int room = (STACK_CAP - STACK_RESERVE - sp) / 7;
            const int k = min(min(32, sp), max(room, 1));
            const bool has = lane < k;
            NodeRec nd;
            nd.cm = make_float4(0.f, 0.f, 0.f, 0.f);
            nd.first_child = 0;
            nd.begin = 0;
            nd.count = 0;
            nd.nchild_level = 0;
            uint32_t id = 0;
            if (has) {
                id = stack[sp - 1 - lane];
                const uint4 *q = reinterpret_cast<const uint4 *>(a.nodes + id);
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
                nd.cm = make_float4(__uint_as_float(q0.x), __uint_as_float(q0.y),
                                    __uint_as_float(q0.z), __uint_as_float(q0.w));
                nd.first_child = q1.x;
                nd.nchild_level = q1.y;
                nd.begin = q1.z;
                nd.count = q1.w;
            }
            sp -= k;
            __syncwarp();

            bool open = false;
            if (has) {
                const float ddx = fmaxf(fabsf(nd.cm.x - cx) - hx, 0.f);
                const float ddy = fmaxf(fabsf(nd.cm.y - cy) - hy, 0.f);
                const float ddz = fmaxf(fabsf(nd.cm.z - cz) - hz, 0.f);
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                const int level = (int)(nd.nchild_level >> 8);
                const float w = ext * __int_as_float((127 - level) << 23);
                open = a.theta2 * d2 < w * w;
            }
            const uint32_t nc = nd.nchild_level & 0xffu;
            const bool open_internal = has && open && nc > 0;
            const bool open_leaf = has && open && nc == 0;
            const bool accept = has && !open;

            const int c_child = open_internal ? (int)nc : 0;
            const int c_leaf = open_leaf ? (int)nd.count : 0;
            int child_incl = c_child, leaf_incl = c_leaf;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, child_incl, o);
                const int u = __shfl_up_sync(FULL, leaf_incl, o);
                if (lane >= o) {
                    child_incl += v;
                    leaf_incl += u;
                }
            }
            {
                const int total = __shfl_sync(FULL, child_incl, 31);
                const int base = sp + child_incl - c_child;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < c_child) stack[base + j] = nd.first_child + j;
                sp += total;
            }

            {  // accepted nodes -> ring (their double-precision {com, mass})
                const unsigned m = __ballot_sync(FULL, accept);
                if (m) {
                    if (accept)
                        list[(head + fill + __popc(m & ((1u << lane) - 1))) & (LIST_CAP - 1)] = x.cm64[id];
                    fill += __popc(m);
                    flush_full();
                }
            }

            {  // particles of opened leaves -> ring, 32 per round
                const int incl = leaf_incl;
                const int total = __shfl_sync(FULL, incl, 31);
                const int excl = incl - c_leaf;
                for (int base = 0; base < total; base += 32) {
                    const int f = base + lane;
                    int owner = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(FULL, incl, (owner + step - 1) & 31);
                        if (v <= f) owner += step;
                    }
                    owner = min(owner, 31);
                    const uint32_t ob = __shfl_sync(FULL, nd.begin, owner);
                    const int oe = __shfl_sync(FULL, excl, owner);
                    if (f < total) list[(head + fill + lane) & (LIST_CAP - 1)] = x.src64[ob + (f - oe)];
                    fill += min(32, total - base);
                    flush_full();
                }
            }
            __syncwarp();
        }
        if (fill > 0) {
            __syncwarp();
            for (int q = slice; q < fill; q += slices)
                eval_entry64(list[(head + q) & (LIST_CAP - 1)], px, py, pz, x.eps2, ax, ay, az);
        }
        for (int o = gpad; o < 32; o <<= 1) {  // combine the slices of each target
            ax += __shfl_xor_sync(FULL, ax, o);
            ay += __shfl_xor_sync(FULL, ay, o);
            az += __shfl_xor_sync(FULL, az, o);
        }
        if (slice == 0 && tl < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ti] : (uint32_t)ti;
            double *o = x.out + (size_t)row * a.dim;
            o[0] = ax;
            o[1] = ay;
            if (a.dim == 3) o[2] = az;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host side.
template <int DIM>
static int sort_by_key(pcuda_ctx *ctx, const float *d_pos, int stride, size_t n, const Frame *d_frame,
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
static void tree_reset(pcuda_ctx *ctx, pcuda_tree *t, size_t n) {
    t->dim = DIM;
    t->bits = Dims<DIM>::BITS;
    t->n = n;
    t->n_nodes = 0;
    t->n_levels = 0;
    t->leaf_size = ctx->leaf_size;
    t->level_begin.clear();
    t->frame = Frame{};
}

// K2: root cube of `n` particle rows -> t->d_frame (device).
template <int DIM>
static int build_frame(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n) {
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
static int build_levels(pcuda_ctx *ctx, pcuda_tree *t, size_t n);

template <int DIM>
static int build(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n,
                 bool keys_only = false) {
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
static int build_levels(pcuda_ctx *ctx, pcuda_tree *t, size_t n) {
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

static int build_dim(pcuda_ctx *ctx, pcuda_tree *t, uint32_t dim, const float *d_particles, size_t n,
                     bool keys_only = false) {
    if (dim == 3) return build<3>(ctx, t, d_particles, n, keys_only);
    if (dim == 2) return build<2>(ctx, t, d_particles, n, keys_only);
    return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
}

static int g_variant = 0;     // experimental traversal variants (tuning hook)
// Tree over double-precision particles: the f32 structure over the rounded records, then the f64
// layer (sources in key order, {com, mass} per node from the f64 positions, bottom-up).
template <int DIM>
static int build64(pcuda_ctx *ctx, pcuda_tree *t, const double *d_particles64, size_t n) {
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

static bool g_count = true;  // instrumentation of the traversal (pcuda_tree_last_counters)
static int g_seg_max = 256;  // largest cell (in targets) that is cut into groups (tuning hook)
static int g_tpl = 2;        // targets per lane in the traversal: 1 (groups of 32) or 2 (groups of 64)

// d_tgt == nullptr: the targets are the tree's own particles (the `&[P]` storage).
// tgt_stride: floats per target row (0 = bare positions, i.e. `dim`).
struct Ext64;
// A forest of trees over the same root cube stored back to back (partitioned build): node and
// source arrays that replace the tree's own, and the roots the walk starts from.
struct ForestView {
    const NodeRec *nodes;
    const float4 *src;
    const uint32_t *d_roots;  // device array
    uint32_t n_roots;
};
static int traverse_sorted(pcuda_ctx *ctx, const pcuda_tree *t, const float4 *tgt_sorted,
                           const uint64_t *tgt_keys, const uint32_t *tgt_perm, size_t na, float theta,
                           float eps, float *d_out, const Ext64 *x64 = nullptr,
                           const ForestView *fv = nullptr);

// Double precision (tree built by build64): d_tgt64 / d_out64 replace d_tgt / d_out; the f32 copy of
// separate targets that keys them is made here.
static int traverse(pcuda_ctx *ctx, const pcuda_tree *t, const float *d_tgt, size_t na, float theta,
                    float eps, float *d_out, int tgt_stride = 0, const double *d_tgt64 = nullptr,
                    double *d_out64 = nullptr, double eps64 = 0.0) {
    const int dim = t->dim;
    const int ts = tgt_stride ? tgt_stride : dim;
    const bool f64 = d_out64 != nullptr;
    if (na == 0) return PCUDA_OK;
    if (t->n == 0) {
        if (f64) PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_out64, 0, na * dim * sizeof(double), ctx->stream));
        else PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, na * dim * sizeof(float), ctx->stream));
        return PCUDA_OK;
    }
    if (f64 && d_tgt64) {  // f32 copy of the target rows, only to key and group them
        PCUDA_CUDA_TRY(ctx, ctx->d_misc.ensure(na * ts * sizeof(float)));
        narrow_kernel<<<(unsigned)std::min<size_t>((na * ts + 255) / 256, 65535), 256, 0, ctx->stream>>>(
            d_tgt64, na * ts, ctx->d_misc.as<float>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        d_tgt = ctx->d_misc.as<float>();
    }
    if (na > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    if (!d_tgt && na != t->n)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "affected == NULL but n_affected != tree size");
    cudaStream_t st = ctx->stream;
    const float4 *tgt_sorted;
    const uint32_t *tgt_perm;
    const uint64_t *tgt_keys;
    if (!d_tgt) {
        tgt_sorted = t->sorted.as<float4>();
        tgt_perm = t->d_perm();
        tgt_keys = t->d_keys();
    } else {
        // key the targets in the tree's frame and process them in key order (coherent groups)
        DevBuf keys[2] = {ctx->d_tgt_keys, ctx->d_tgt_keys_alt};
        DevBuf perm[2] = {ctx->d_tgt_perm, ctx->d_tgt_perm_alt};
        int cur = 0;
        int s = dim == 3 ? sort_by_key<3>(ctx, d_tgt, ts, na, t->d_frame.as<Frame>(), keys, perm, &cur, ctx->d_cub_tmp)
                         : sort_by_key<2>(ctx, d_tgt, ts, na, t->d_frame.as<Frame>(), keys, perm, &cur, ctx->d_cub_tmp);
        ctx->d_tgt_keys = keys[0];
        ctx->d_tgt_keys_alt = keys[1];
        ctx->d_tgt_perm = perm[0];
        ctx->d_tgt_perm_alt = perm[1];
        PCUDA_TRY(s);
        PCUDA_CUDA_TRY(ctx, ctx->d_tgt_sorted.ensure(na * (f64 ? sizeof(double4) : sizeof(float4))));
        const uint32_t *p = perm[cur].as<uint32_t>();
        if (f64 && dim == 3)
            gather64_kernel<3><<<(unsigned)((na + 255) / 256), 256, 0, st>>>(d_tgt64, ts, false, (int)na, p,
                                                                             ctx->d_tgt_sorted.as<double4>());
        else if (f64)
            gather64_kernel<2><<<(unsigned)((na + 255) / 256), 256, 0, st>>>(d_tgt64, ts, false, (int)na, p,
                                                                             ctx->d_tgt_sorted.as<double4>());
        else if (dim == 3)
            gather_kernel<3><<<(unsigned)((na + 255) / 256), 256, 0, st>>>(d_tgt, ts, false, (int)na, p,
                                                                           ctx->d_tgt_sorted.as<float4>());
        else
            gather_kernel<2><<<(unsigned)((na + 255) / 256), 256, 0, st>>>(d_tgt, ts, false, (int)na, p,
                                                                           ctx->d_tgt_sorted.as<float4>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        tgt_sorted = ctx->d_tgt_sorted.as<float4>();
        tgt_perm = p;
        tgt_keys = keys[cur].as<uint64_t>();
    }
    if (f64) {
        Ext64 x;
        x.src64 = t->sorted64.as<double4>();
        x.cm64 = t->moments.as<double4>();
        x.tgt64 = d_tgt64 ? ctx->d_tgt_sorted.as<double4>() : t->sorted64.as<double4>();
        x.out = d_out64;
        x.eps2 = eps64 * eps64;
        return traverse_sorted(ctx, t, nullptr, tgt_keys, tgt_perm, na, theta, eps, nullptr, &x);
    }
    return traverse_sorted(ctx, t, tgt_sorted, tgt_keys, tgt_perm, na, theta, eps, d_out);
}

// Targets already in key order ({x,y,z,_} records + their keys in the tree's frame); tgt_perm maps
// traversal order to the output row (nullptr: out row = traversal position).
static int traverse_sorted(pcuda_ctx *ctx, const pcuda_tree *t, const float4 *tgt_sorted,
                           const uint64_t *tgt_keys, const uint32_t *tgt_perm, size_t na, float theta,
                           float eps, float *d_out, const Ext64 *x64, const ForestView *fv) {
    const int dim = t->dim;
    if (fv && (x64 || t->order == 2 || g_tpl != 2 || g_variant))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "a forest is walked by traverse2_kernel only");
    cudaStream_t st = ctx->stream;
    const int group_cap = x64 ? 32 : 32 * g_tpl;  // the f64 walk holds one target per lane
    // K5a: groups from the target keys
    const int n = (int)na;
    PCUDA_CUDA_TRY(ctx, ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, 8 * sizeof(unsigned long long), st));
    uint32_t *d_work = reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 3);
    uint32_t *d_ngroups = reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 4);
    // d_stack layout: L (n bytes, padded) | flag (n u32) | pos (n u32) | group_start (n + 1 u32) |
    // hard-boundary bits (one word per 32 targets, padded to whole blocks)
    const size_t n4 = ((size_t)n + 3) & ~size_t(3);
    const size_t nhw = ((size_t)n + GROUP_BLOCK - 1) / GROUP_BLOCK * (GROUP_BLOCK / 32);
    PCUDA_CUDA_TRY(ctx, ctx->d_stack.ensure(n4 + (3 * (size_t)n + 1 + nhw) * 4));
    uint8_t *d_L = ctx->d_stack.as<uint8_t>();
    uint32_t *d_flag = reinterpret_cast<uint32_t *>(d_L + n4);
    uint32_t *d_pos = d_flag + n;
    uint32_t *d_gstart = d_pos + n;
    uint32_t *d_hard = d_gstart + n + 1;
    const unsigned nb256 = (unsigned)((n + 255) / 256);
    if (dim == 3) boundary_levels<3><<<nb256, 256, 0, st>>>(tgt_keys, n, d_L);
    else boundary_levels<2><<<nb256, 256, 0, st>>>(tgt_keys, n, d_L);
    const unsigned ngb = (unsigned)((n + GROUP_BLOCK - 1) / GROUP_BLOCK);
    hard_flags<<<ngb, GROUP_BLOCK, 0, st>>>(d_L, n, t->bits, g_seg_max, d_hard);
    group_flags<<<ngb, GROUP_BLOCK, 0, st>>>(d_hard, n, g_seg_max, group_cap, d_flag);
    size_t tmp = 0;
    PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_flag, d_pos, n, st));
    PCUDA_CUDA_TRY(ctx, ctx->d_cub_tmp.ensure(tmp));
    PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, tmp, d_flag, d_pos, n, st));
    scatter_groups<<<nb256, 256, 0, st>>>(d_flag, d_pos, n, d_gstart, d_ngroups);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 6;

    TravArgs a;
    a.nodes = t->nodes.as<NodeRec>();
    a.src = t->sorted.as<float4>();
    a.tgt = tgt_sorted;
    a.tgt_perm = tgt_perm;
    a.group_start = d_gstart;
    a.n_groups = d_ngroups;
    a.work = d_work;
    a.out = d_out;
    a.counters = ctx->d_counters.as<unsigned long long>();
    a.n_tgt = n;
    a.dim = dim;
    a.frame = t->d_frame.as<Frame>();
    a.theta2 = theta * theta;
    a.eps2 = eps * eps;
    a.n_roots = 1;
    a.roots = nullptr;
    if (fv) {
        a.nodes = fv->nodes;
        a.src = fv->src;
        a.n_roots = fv->n_roots;
        a.roots = fv->d_roots;
    }
    const size_t max_groups = ((size_t)n + 7) / 8;  // enough warps for small inputs, persistent beyond
    const unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4,
                                                       (max_groups + TRAV_WARPS - 1) / TRAV_WARPS);
    if (!x64 && t->order == 2) {
        const unsigned blocksq = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 6,
                                                           (max_groups + TRAVQ_WARPS - 1) / TRAVQ_WARPS);
        traverse2q_kernel<<<blocksq, TRAVQ_WARPS * 32, 0, st>>>(a, t->quad.as<float4>());
    } else if (x64) {
        const unsigned blocks64 = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 8,
                                                            (max_groups + TRAV64_WARPS - 1) / TRAV64_WARPS);
        traverse64_kernel<<<blocks64, TRAV64_WARPS * 32, 0, st>>>(a, *x64);
    } else if (g_tpl == 2 && g_variant) {
        if (g_variant == 1) traverse2_kernel<false, 1><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else if (g_variant == 2) traverse2_kernel<false, 2><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else traverse2_kernel<false, 3><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
    } else if (g_tpl == 2 && fv) {
        if (g_count) traverse2_kernel<true, 0, true><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else traverse2_kernel<false, 0, true><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
    } else if (g_tpl == 2) {
        if (g_count) traverse2_kernel<true><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else traverse2_kernel<false><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
    } else {
        if (g_count) traverse_kernel<true><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else traverse_kernel<false><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
    }
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return PCUDA_OK;
}

static int read_counters(pcuda_ctx *ctx) {
    if (!ctx->d_counters.p) return PCUDA_OK;
    unsigned long long h[6];  // [3], [4] hold the work dispenser and the group count
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_counters.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) ctx->last_counters[i] = h[i];
    ctx->last_counters[3] = h[5];
    ctx->last_counters[4] = h[4] & 0xffffffffull;  // number of target groups
    return PCUDA_OK;
}

// One-shot Barnes-Hut with device pointers: build over `affecting`, traverse for `affected`.
static int oneshot_dev(pcuda_ctx *ctx, uint32_t dim, const float *d_aff, size_t na, const float *d_src,
                       size_t nb, float theta, float eps, float *d_out, int tgt_stride = 0) {
    if (!d_aff && na != nb)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "affected == NULL means affected == affecting, but n_affected != n_affecting");
    if (!ctx->call_tree) ctx->call_tree = new pcuda_tree();
    phase_begin(ctx, PH_BUILD);
    PCUDA_TRY(build_dim(ctx, ctx->call_tree, dim, d_src, nb));
    phase_end(ctx, PH_BUILD);
    phase_begin(ctx, PH_COMPUTE);
    PCUDA_TRY(traverse(ctx, ctx->call_tree, d_aff, na, theta, eps, d_out, tgt_stride));
    phase_end(ctx, PH_COMPUTE);
    return PCUDA_OK;
}

static int oneshot_dev64(pcuda_ctx *ctx, uint32_t dim, const double *d_aff, size_t na, const double *d_src,
                         size_t nb, double theta, double eps, double *d_out, int tgt_stride = 0) {
    if (!d_aff && na != nb)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "affected == NULL means affected == affecting, but n_affected != n_affecting");
    if (dim != 2 && dim != 3) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    if (!ctx->call_tree) ctx->call_tree = new pcuda_tree();
    phase_begin(ctx, PH_BUILD);
    PCUDA_TRY(dim == 3 ? build64<3>(ctx, ctx->call_tree, d_src, nb) : build64<2>(ctx, ctx->call_tree, d_src, nb));
    phase_end(ctx, PH_BUILD);
    phase_begin(ctx, PH_COMPUTE);
    PCUDA_TRY(traverse(ctx, ctx->call_tree, nullptr, na, (float)theta, 0.f, nullptr, tgt_stride, d_aff, d_out,
                       eps));
    phase_end(ctx, PH_COMPUTE);
    return PCUDA_OK;
}

static int oneshot_host64(pcuda_ctx *ctx, uint32_t dim, const double *aff, size_t na, const double *src,
                          size_t nb, double theta, double eps, double *out) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if ((na && !out) || (nb && !src))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (!aff && na != nb)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "affected == NULL means affected == affecting, but n_affected != n_affecting");
    if (na > 0x7fffffffull || nb > 0x7fffffffull)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (na == 0) return PCUDA_OK;
    const size_t src_bytes = nb * (dim + 1) * sizeof(double), tgt_bytes = na * dim * sizeof(double);
    phase_begin(ctx, PH_UPLOAD);
    double *d_src = nullptr, *d_tgt = nullptr;
    if (nb) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(src_bytes));
        d_src = ctx->d_affecting.as<double>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (aff) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affected.ensure(tgt_bytes));
        d_tgt = ctx->d_affected.as<double>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_tgt, aff, tgt_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(tgt_bytes));
    phase_end(ctx, PH_UPLOAD);
    PCUDA_TRY(oneshot_dev64(ctx, dim, d_tgt, na, d_src, nb, theta, eps, ctx->d_out.as<double>()));
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out.p, tgt_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

// Multi-GPU step (one process per GPU), "replicated build": every rank owns the contiguous block
// [rank * cap, rank * cap + n_local) of the n_total particles (cap = ceil(n_total / world)).  The
// local records are all-gathered in place over NVLink and every GPU builds the identical tree over
// all n_total particles.  The traversal is sharded by KEY RANGE, not by input block: rank r walks
// the tree for the sorted particles [r * cap, (r + 1) * cap) — spatially compact, so its target
// groups are as tight as on one GPU and alias the tree's own sorted records (no target sort) —
// writes their accelerations in key order, the per-range results are all-gathered in place
// (12 B per particle), and each rank picks the rows of the particles it owns through the sort
// permutation.  (Sharding the traversal by input block made every rank walk a sparse random sample
// of the cloud: 8.7 ms instead of 6.0 ms per rank at N = 10M on 4 GPUs.)
__global__ void __launch_bounds__(256) pick_owned_rows(const float *__restrict__ acc_sorted,
                                                       const uint32_t *__restrict__ perm, int n,
                                                       uint32_t lo, uint32_t hi,
                                                       float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t orig = perm[i];
    if (orig < lo || orig >= hi) return;
    float *o = out + (size_t)(orig - lo) * 3;
    o[0] = acc_sorted[(size_t)i * 3 + 0];
    o[1] = acc_sorted[(size_t)i * 3 + 1];
    o[2] = acc_sorted[(size_t)i * 3 + 2];
}

// Routing of the per-range accelerations to the ranks that own the particles.  The all-gather
// above moves 12 B x N to every rank although a rank needs only the rows of its own block; with
// ncclSend / ncclRecv available each row (acceleration + original index, 16 B) is sent to its owner
// only.  Before the traversal: owners counted per row, counts all-gathered (one synchronisation),
// every row given a slot in an owner-bucketed send buffer; the traversal then writes straight into
// that buffer (its row map is `pos`), and one variable all-to-all plus a scatter finish the step.
struct OwnerOffsets {
    uint32_t off[MAX_PARTS];
};

__global__ void __launch_bounds__(256) owner_hist(const uint32_t *__restrict__ idx, int n, uint32_t cap,
                                                  uint32_t *__restrict__ cnt) {
    __shared__ uint32_t s_cnt[MAX_PARTS];
    if (threadIdx.x < MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&s_cnt[idx[i] / cap], 1u);
    __syncthreads();
    if (threadIdx.x < MAX_PARTS && s_cnt[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256) owner_positions(const uint32_t *__restrict__ idx, int n, uint32_t cap,
                                                       OwnerOffsets send_off, uint32_t *__restrict__ cursor,
                                                       uint32_t *__restrict__ pos,
                                                       uint32_t *__restrict__ idx_send) {
    __shared__ uint32_t s_cnt[MAX_PARTS], s_base[MAX_PARTS];
    if (threadIdx.x < MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t o = 0, mine = 0, orig = 0;
    if (i < n) {
        orig = idx[i];
        o = orig / cap;
        mine = atomicAdd(&s_cnt[o], 1u);
    }
    __syncthreads();
    if (threadIdx.x < MAX_PARTS && s_cnt[threadIdx.x])
        s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (i < n) {
        const uint32_t p = send_off.off[o] + s_base[o] + mine;
        pos[i] = p;
        idx_send[p] = orig;
    }
}

__global__ void __launch_bounds__(256) scatter_rows(const float *__restrict__ acc,
                                                    const uint32_t *__restrict__ idx, int n, uint32_t lo,
                                                    float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *o = out + (size_t)(idx[i] - lo) * 3;
    o[0] = acc[(size_t)i * 3 + 0];
    o[1] = acc[(size_t)i * 3 + 1];
    o[2] = acc[(size_t)i * 3 + 2];
}

struct RoutePlan {
    size_t send_off[MAX_PARTS], send_cnt[MAX_PARTS], recv_off[MAX_PARTS], recv_cnt[MAX_PARTS];
    size_t n_rows = 0, n_recv = 0;
    uint32_t *d_pos = nullptr;
    float *d_acc_send = nullptr;
};

// ------------------------------------------------------------------------------------------------
// Key-range-partitioned build.  The replicated build costs every GPU the whole
// sort + tree (2.3 ms at N = 10M) however many GPUs share the traversal.  Here the key space is cut
// into `parts` ranges of about equal population and every part builds the tree of ITS particles
// only — over the same root cube, with the same level / leaf rules.  The per-part trees are stored
// back to back and joined by a small TOP TREE:
//
//   1. keys of all particles in the common frame (replicated: 0.1 ms at N = 10M);
//   2. splitters = quantiles of a regular sample of <= 65536 keys (sorted by every rank alike),
//      per-part populations counted in one pass;
//   3. stable selection of the part's (key, index) pairs, sort, gather, level-wise build: all over
//      n / parts particles;
//   4. exchange: node records and sort permutations are all-gathered into equal slots
//      (child / particle indices rebased to the slot), the sources are re-gathered locally from the
//      raw records that every rank already holds (cheaper than sending them once more);
//   5. cells that straddle a range boundary exist in several parts as PARTIAL cells (each with the
//      moments of its own particles).  On every level of a part only the first and the last node
//      can be partial (nodes of a level are in key order), so at most 2 x 22 x parts cells are
//      involved: their records, key prefixes, double-precision moments and children are brought to
//      the host, partial cells with the same (level, prefix) are merged — moments added in part
//      order, children = the complete children of every part plus the merged children — and the
//      merged cells are appended to the node array as the top tree.  Where a part's share of a
//      merged cell is a LEAF (<= leaf_size of the part's particles) that leaf becomes one more child
//      of the merged cell, with the cell's own level (its particles may lie anywhere in the cell);
//      a merged cell with more than 8 children keeps 7 and links the others behind a continuation
//      node of its own level.  A walk from the top root meets every particle exactly once and sees
//      the same cells, with the same centres of mass (up to the order of the f64 additions), as a
//      walk of the single tree.  (Walking the per-part trees as a plain forest, partial cells and all, is
//      also exact at theta = 0 but less accurate at theta > 0 — a half-empty cell has a large
//      quadrupole: median error 6.8e-4 instead of 2.6e-4 at N = 2M, 8 parts.)
//   6. every rank walks the joined tree for the targets of its own key range.
struct PartRange {
    const uint64_t *keys;
    const uint64_t *split;
    int part;
    __device__ __forceinline__ bool operator()(const uint32_t &i) const {
        const uint64_t k = keys[i];
        return k >= split[part] && (k < split[part + 1] || split[part + 1] == ~0ull);
    }
};

__global__ void __launch_bounds__(256) sample_keys(const uint64_t *__restrict__ keys, size_t stride,
                                                   int m, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = keys[(size_t)j * stride];
}

// split[0] = 0, split[q] = q-th parts-quantile of the sorted sample, split[parts] = ~0 (inclusive).
__global__ void pick_splitters(const uint64_t *__restrict__ sorted_sample, int m, int parts,
                               uint64_t *__restrict__ split, uint32_t *__restrict__ counts) {
    const int q = threadIdx.x;
    if (q <= parts) {
        split[q] = q == 0 ? 0ull : q == parts ? ~0ull : sorted_sample[(size_t)q * m / parts];
        counts[q] = 0;
    }
}

__global__ void __launch_bounds__(256) count_parts(const uint64_t *__restrict__ keys, int n,
                                                   const uint64_t *__restrict__ split, int parts,
                                                   uint32_t *__restrict__ counts) {
    __shared__ uint32_t s_cnt[MAX_PARTS];
    __shared__ uint64_t s_split[MAX_PARTS + 1];
    if (threadIdx.x < MAX_PARTS) s_cnt[threadIdx.x] = 0;
    if ((int)threadIdx.x <= parts) s_split[threadIdx.x] = split[threadIdx.x];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        int q = 0;
        while (q + 1 < parts && k >= s_split[q + 1]) ++q;
        atomicAdd(&s_cnt[q], 1u);
    }
    __syncthreads();
    if ((int)threadIdx.x < parts && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256) take_keys(const uint64_t *__restrict__ keys,
                                                 const uint32_t *__restrict__ idx, int n,
                                                 uint64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keys[idx[i]];
}

// Local node records -> their slot of the joined array: child links and particle ranges rebased.
__global__ void __launch_bounds__(256) copy_rebase_nodes(const NodeRec *__restrict__ in, uint32_t n_nodes,
                                                         uint32_t node_base, uint32_t part_base,
                                                         NodeRec *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    NodeRec r = in[i];
    if (r.nchild_level & 0xffu) r.first_child += node_base;
    r.begin += part_base;
    out[i] = r;
}

constexpr uint32_t NO_PARTICLE = 0xffffffffu;  // padding of a permutation slot

__global__ void __launch_bounds__(256) copy_pad_perm(const uint32_t *__restrict__ in, uint32_t n,
                                                     uint32_t slot, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slot) out[i] = i < n ? in[i] : NO_PARTICLE;
}

// Sources of all parts in slot order, from the raw {x,y,z,mu} rows and the permutation slots.
__global__ void __launch_bounds__(256) gather_forest(const float4 *__restrict__ raw,
                                                     const uint32_t *__restrict__ perm, size_t n_slots,
                                                     float4 *__restrict__ sorted) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const uint32_t o = perm[i];
    if (o != NO_PARTICLE) sorted[i] = raw[o];
}

// What a part tells the others about its tree besides the node records: the level table and, for
// the first and the last node of every level (the only possibly partial cells), the key prefix of
// the cell and its double-precision moments {sum m x, sum m y, sum m z, sum m}.
constexpr int TOP_LEVELS = Dims<3>::BITS + 1;  // 22
struct PartPack {
    uint32_t n_nodes, n_levels;
    uint32_t level_begin[TOP_LEVELS + 2];
    uint64_t prefix[TOP_LEVELS][2];
    double mom[TOP_LEVELS][2][4];
};

__global__ void fill_pack(const NodeRec *__restrict__ nodes, const double *__restrict__ mom,
                          const uint64_t *__restrict__ keys, const BuildState *__restrict__ st,
                          uint32_t n_nodes, uint32_t n_levels, PartPack *__restrict__ out) {
    const int t = threadIdx.x;
    if (t == 0) {
        out->n_nodes = n_nodes;
        out->n_levels = n_levels;
    }
    if (t < TOP_LEVELS + 2) out->level_begin[t] = n_nodes ? st->level_begin[t] : 0u;
    if (t < 2 * TOP_LEVELS) {
        const int l = t >> 1, side = t & 1;
        uint64_t pre = 0;
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        if (n_nodes && l < (int)n_levels) {
            const uint32_t idx = side ? st->level_begin[l + 1] - 1 : st->level_begin[l];
            pre = keys[nodes[idx].begin] >> (3 * (Dims<3>::BITS - l));
            for (int c = 0; c < 4; ++c) m[c] = mom[(size_t)idx * 4 + c];
        }
        out->prefix[l][side] = pre;
        for (int c = 0; c < 4; ++c) out->mom[l][side][c] = m[c];
    }
}

// Boundary nodes of every part and their children, from the joined (rebased) node array.
struct BoundaryRec {
    NodeRec node;
    NodeRec child[8];
};
struct PartBases {
    uint32_t node_base[MAX_PARTS];
};

__global__ void __launch_bounds__(2 * TOP_LEVELS * 9) collect_boundary(const NodeRec *__restrict__ nodes,
                                                                       const PartPack *__restrict__ packs,
                                                                       PartBases bases,
                                                                       BoundaryRec *__restrict__ out) {
    const int q = blockIdx.x;
    const int t = threadIdx.x / 9, j = threadIdx.x % 9;  // t = (level, side), j = 0: node, 1..8: child
    const int l = t >> 1, side = t & 1;
    const PartPack &pk = packs[q];
    if (pk.n_nodes == 0 || l >= (int)pk.n_levels) return;
    const uint32_t local = side ? pk.level_begin[l + 1] - 1 : pk.level_begin[l];
    const NodeRec nd = nodes[bases.node_base[q] + local];
    BoundaryRec *o = out + ((size_t)q * TOP_LEVELS + l) * 2 + side;
    if (j == 0) o->node = nd;
    else if (j - 1 < (int)(nd.nchild_level & 0xffu)) o->child[j - 1] = nodes[nd.first_child + j - 1];
}

}  // namespace bh
}  // namespace pcuda

struct pcuda_forest {
    pcuda_tree *local = nullptr;       // tree of this rank's (or the current part's) key range
    pcuda::DevBuf gkeys, gidx;         // keys of ALL particles in input order (+ identity scratch)
    pcuda::DevBuf sample[2], split, counts, sel_tmp, sel_count;
    pcuda::DevBuf nodes, sorted, perm, keys, acc;  // the joined tree: equal slots per part (+ top tree)
    pcuda::DevBuf packs, stage, roots;
    pcuda::DevBuf route_cnt, route_pos, route_idx_send, route_acc_send, route_idx_recv, route_acc_recv;
    uint32_t *h_route = nullptr;                   // pinned: world x MAX_PARTS owner counts
    pcuda::bh::PartPack *h_packs = nullptr;        // pinned
    pcuda::bh::BoundaryRec *h_stage = nullptr;     // pinned
    cudaEvent_t ev_stage = nullptr;
};

namespace pcuda {

void forest_free(pcuda_ctx *ctx) {
    pcuda_forest *f = ctx->forest;
    if (!f) return;
    if (f->local) tree_free(ctx, f->local);
    DevBuf *bufs[] = {&f->gkeys, &f->gidx, &f->sample[0], &f->sample[1], &f->split, &f->counts,
                      &f->sel_tmp, &f->sel_count, &f->nodes, &f->sorted, &f->perm, &f->keys, &f->acc,
                      &f->packs, &f->stage, &f->roots, &f->route_cnt, &f->route_pos, &f->route_idx_send,
                      &f->route_acc_send, &f->route_idx_recv, &f->route_acc_recv};
    for (DevBuf *b : bufs) b->release();
    if (f->h_route) cudaFreeHost(f->h_route);
    if (f->h_packs) cudaFreeHost(f->h_packs);
    if (f->h_stage) cudaFreeHost(f->h_stage);
    if (f->ev_stage) cudaEventDestroy(f->ev_stage);
    delete f;
    ctx->forest = nullptr;
}

namespace bh {

constexpr size_t TOP_CAP = 4096;  // top-tree nodes: <= 1 + 8 * 22 * MAX_PARTS

static int forest_of(pcuda_ctx *ctx, pcuda_forest **out) {
    if (!ctx->forest) {
        pcuda_forest *f = new pcuda_forest();
        f->local = new pcuda_tree();
        ctx->forest = f;
    }
    pcuda_forest *f = ctx->forest;
    if (!f->h_packs) PCUDA_CUDA_TRY(ctx, cudaHostAlloc((void **)&f->h_packs, MAX_PARTS * sizeof(PartPack), cudaHostAllocDefault));
    if (!f->h_stage)
        PCUDA_CUDA_TRY(ctx, cudaHostAlloc((void **)&f->h_stage, MAX_PARTS * TOP_LEVELS * 2 * sizeof(BoundaryRec),
                                          cudaHostAllocDefault));
    if (!f->ev_stage) PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&f->ev_stage, cudaEventDisableTiming));
    if (!f->h_route)
        PCUDA_CUDA_TRY(ctx, cudaHostAlloc((void **)&f->h_route, MAX_PARTS * MAX_PARTS * sizeof(uint32_t),
                                          cudaHostAllocDefault));
    PCUDA_CUDA_TRY(ctx, f->route_cnt.ensure((MAX_PARTS * MAX_PARTS + MAX_PARTS) * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->packs.ensure(MAX_PARTS * sizeof(PartPack)));
    PCUDA_CUDA_TRY(ctx, f->stage.ensure(MAX_PARTS * TOP_LEVELS * 2 * sizeof(BoundaryRec)));
    PCUDA_CUDA_TRY(ctx, f->roots.ensure(MAX_ROOTS * sizeof(uint32_t)));
    *out = f;
    return PCUDA_OK;
}

// Steps 1-2: frame, keys, splitters, populations (host copy in counts_h).  One synchronisation.
static int forest_partition(pcuda_ctx *ctx, pcuda_forest *f, const float *d_particles, size_t n,
                            int parts, uint32_t counts_h[MAX_PARTS]) {
    cudaStream_t st = ctx->stream;
    pcuda_tree *t = f->local;
    PCUDA_TRY(build_frame<3>(ctx, t, d_particles, n));
    PCUDA_CUDA_TRY(ctx, f->gkeys.ensure(n * sizeof(uint64_t)));
    PCUDA_CUDA_TRY(ctx, f->gidx.ensure(n * sizeof(uint32_t)));
    encode_kernel<3><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        d_particles, 4, (int)n, t->d_frame.as<Frame>(), f->gkeys.as<uint64_t>(), f->gidx.as<uint32_t>());
    const int m = (int)std::min<size_t>(n, 65536);
    const size_t stride = n / (size_t)m;
    for (int i = 0; i < 2; ++i) PCUDA_CUDA_TRY(ctx, f->sample[i].ensure((size_t)m * sizeof(uint64_t)));
    PCUDA_CUDA_TRY(ctx, f->split.ensure((MAX_PARTS + 1) * sizeof(uint64_t)));
    PCUDA_CUDA_TRY(ctx, f->counts.ensure((MAX_PARTS + 1) * sizeof(uint32_t)));
    sample_keys<<<(m + 255) / 256, 256, 0, st>>>(f->gkeys.as<uint64_t>(), stride, m,
                                                 f->sample[0].as<uint64_t>());
    cub::DoubleBuffer<uint64_t> sb(f->sample[0].as<uint64_t>(), f->sample[1].as<uint64_t>());
    size_t tmp = 0;
    PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp, sb, m, 0, 63, st));
    PCUDA_CUDA_TRY(ctx, f->sel_tmp.ensure(tmp));
    PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(f->sel_tmp.p, tmp, sb, m, 0, 63, st));
    pick_splitters<<<1, 32, 0, st>>>(sb.Current(), m, parts, f->split.as<uint64_t>(),
                                     f->counts.as<uint32_t>());
    count_parts<<<(unsigned)std::min<size_t>((size_t)ctx->sm_count * 8, (n + 255) / 256), 256, 0, st>>>(
        f->gkeys.as<uint64_t>(), (int)n, f->split.as<uint64_t>(), parts, f->counts.as<uint32_t>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 5 + 9;
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(counts_h, f->counts.p, parts * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return PCUDA_OK;
}

// Step 3 for part q (population `count`): f->local becomes the tree of the part's particles and
// the part's pack is written to d_pack.  `slot` >= count: capacity of the key / permutation buffers.
static int forest_build_part(pcuda_ctx *ctx, pcuda_forest *f, const float *d_particles, size_t n, int q,
                             size_t count, size_t slot, PartPack *d_pack) {
    cudaStream_t st = ctx->stream;
    pcuda_tree *t = f->local;
    tree_reset<3>(ctx, t, count);
    PCUDA_CUDA_TRY(ctx, t->scan_in.ensure(sizeof(BuildState)));
    if (count) {
        for (int i = 0; i < 2; ++i) {
            PCUDA_CUDA_TRY(ctx, t->keys[i].ensure(slot * sizeof(uint64_t)));
            PCUDA_CUDA_TRY(ctx, t->perm[i].ensure(slot * sizeof(uint32_t)));
        }
        PCUDA_CUDA_TRY(ctx, f->sel_count.ensure(sizeof(uint32_t)));
        PartRange in_part{f->gkeys.as<uint64_t>(), f->split.as<uint64_t>(), q};
        cub::CountingInputIterator<uint32_t> all(0u);
        size_t tmp = 0;
        PCUDA_CUDA_TRY(ctx, cub::DeviceSelect::If(nullptr, tmp, all, t->perm[0].as<uint32_t>(),
                                                  f->sel_count.as<uint32_t>(), (int)n, in_part, st));
        PCUDA_CUDA_TRY(ctx, f->sel_tmp.ensure(tmp));
        PCUDA_CUDA_TRY(ctx, cub::DeviceSelect::If(f->sel_tmp.p, tmp, all, t->perm[0].as<uint32_t>(),
                                                  f->sel_count.as<uint32_t>(), (int)n, in_part, st));
        take_keys<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(
            f->gkeys.as<uint64_t>(), t->perm[0].as<uint32_t>(), (int)count, t->keys[0].as<uint64_t>());
        cub::DoubleBuffer<uint64_t> kb(t->keys[0].as<uint64_t>(), t->keys[1].as<uint64_t>());
        cub::DoubleBuffer<uint32_t> vb(t->perm[0].as<uint32_t>(), t->perm[1].as<uint32_t>());
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int)count, 0, 63, st));
        PCUDA_CUDA_TRY(ctx, t->cub_tmp.ensure(tmp));
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(t->cub_tmp.p, tmp, kb, vb, (int)count, 0, 63, st));
        t->cur = kb.selector;
        PCUDA_CUDA_TRY(ctx, t->sorted.ensure(count * sizeof(float4)));
        gather_kernel<3><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_particles, 4, true, (int)count,
                                                                          t->d_perm(), t->sorted.as<float4>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 2 + 1 + 9 + 1;
        PCUDA_TRY(build_levels<3>(ctx, t, count));
    }
    fill_pack<<<1, 64, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double>(), count ? t->d_keys() : nullptr,
                                t->scan_in.as<BuildState>(), (uint32_t)t->n_nodes, (uint32_t)t->n_levels, d_pack);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return PCUDA_OK;
}

// Step 5 on the host.  packs / stage: every part's pack and boundary records (stage indexed
// [part][level][side]); node_base: first node of every part in the joined array; top_base: where
// the top tree goes.  Out: the top-tree nodes and the start nodes of the walk.
static int merge_top_tree(pcuda_ctx *ctx, int parts, const PartPack *packs, const BoundaryRec *stage,
                          const uint32_t *node_base, uint32_t top_base, std::vector<NodeRec> &top,
                          std::vector<uint32_t> &roots) {
    struct Inst {
        int q, l, side;
        uint32_t gi;
        const BoundaryRec *b;
    };
    struct Cell {  // a (level, prefix) that occurs as a boundary node
        int l;
        uint64_t prefix;
        std::vector<int> inst;  // indices into `insts`, in part order
    };
    std::vector<Inst> insts;
    std::vector<Cell> cells;
    top.clear();
    roots.clear();
    std::map<std::pair<int, uint64_t>, int> cell_index;  // (level, prefix) -> cells[]
    std::map<uint32_t, int> inst_index;                  // joined node index -> insts[]
    auto find_cell = [&](int l, uint64_t prefix) -> int {
        auto it = cell_index.find({l, prefix});
        return it == cell_index.end() ? -1 : it->second;
    };
    int nonempty = 0, last_nonempty = -1;
    for (int q = 0; q < parts; ++q) {
        const PartPack &pk = packs[q];
        if (pk.n_nodes == 0) continue;
        ++nonempty;
        last_nonempty = q;
        if (pk.n_levels > (uint32_t)TOP_LEVELS) return fail(ctx, PCUDA_ERR_CUDA, "part %d reports %u levels", q, pk.n_levels);
        for (int l = 0; l < (int)pk.n_levels; ++l) {
            const uint32_t lb = pk.level_begin[l], le = pk.level_begin[l + 1];
            for (int side = 0; side < 2; ++side) {
                if (side == 1 && le - lb == 1) continue;  // one node on the level: first == last
                Inst in;
                in.q = q;
                in.l = l;
                in.side = side;
                in.gi = node_base[q] + (side ? le - 1 : lb);
                in.b = stage + ((size_t)q * TOP_LEVELS + l) * 2 + side;
                int c = find_cell(l, pk.prefix[l][side]);
                if (c < 0) {
                    Cell nc;
                    nc.l = l;
                    nc.prefix = pk.prefix[l][side];
                    cells.push_back(nc);
                    c = (int)cells.size() - 1;
                    cell_index[{l, nc.prefix}] = c;
                }
                cells[c].inst.push_back((int)insts.size());
                inst_index[in.gi] = (int)insts.size();
                insts.push_back(in);
            }
        }
    }
    if (nonempty == 0) return PCUDA_OK;
    if (nonempty == 1) {
        roots.push_back(node_base[last_nonempty]);
        return PCUDA_OK;
    }
    auto merged = [&](int c) { return cells[c].inst.size() >= 2; };
    // boundary node -> its cell (to recognise children that are themselves merged)
    auto cell_of_node = [&](uint32_t gi, int l) -> int {
        auto it = inst_index.find(gi);
        if (it == inst_index.end()) return -1;
        const Inst &in = insts[it->second];
        return find_cell(l, packs[in.q].prefix[in.l][in.side]);
    };
    const int root_cell = find_cell(0, 0);
    if (root_cell < 0 || !merged(root_cell)) return fail(ctx, PCUDA_ERR_CUDA, "top tree: the root cell is not shared");
    auto record_of = [&](int c) {  // merged cell: moments added in part order
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        uint32_t count = 0;
        for (int ii : cells[c].inst) {
            const Inst &in = insts[ii];
            for (int k = 0; k < 4; ++k) m[k] += packs[in.q].mom[in.l][in.side][k];
            count += in.b->node.count;
        }
        const NodeRec &first = insts[cells[c].inst[0]].b->node;
        NodeRec r;
        if (m[3] == 0.0) r.cm = make_float4(first.cm.x, first.cm.y, first.cm.z, 0.f);
        else r.cm = make_float4((float)(m[0] / m[3]), (float)(m[1] / m[3]), (float)(m[2] / m[3]), (float)m[3]);
        r.first_child = 0;
        r.nchild_level = (uint32_t)cells[c].l << 8;
        r.begin = first.begin;
        r.count = count;
        return r;
    };
    // A node of the top tree that still needs its children written: a merged cell (cell >= 0) or a
    // continuation node (a merged cell with more than 8 children keeps 7 and links the rest).
    struct Kid {
        NodeRec rec;
        int cell;  // >= 0: merged cell to expand
    };
    struct Pending {
        uint32_t me;
        int level;
        std::vector<Kid> kids;
    };
    auto kids_of_cell = [&](int c) {
        std::vector<Kid> kids;
        std::vector<int> listed;
        for (int ii : cells[c].inst) {
            const Inst &in = insts[ii];
            const uint32_t nc = in.b->node.nchild_level & 0xffu;
            if (nc == 0) {  // this part's share of the cell is a leaf: a child leaf of the cell's own level
                kids.push_back({in.b->node, -1});
                continue;
            }
            for (uint32_t j = 0; j < nc; ++j) {
                const int cc = cell_of_node(in.b->node.first_child + j, in.l + 1);
                if (cc >= 0 && merged(cc)) {
                    bool seen = false;
                    for (int k : listed) seen |= k == cc;
                    if (seen) continue;
                    listed.push_back(cc);
                    kids.push_back({record_of(cc), cc});
                } else {
                    kids.push_back({in.b->child[j], -1});  // complete cell: its subtree stays in its part
                }
            }
        }
        return kids;
    };
    std::vector<Pending> queue;
    top.push_back(record_of(root_cell));
    queue.push_back({0u, 0, kids_of_cell(root_cell)});
    for (size_t h = 0; h < queue.size(); ++h) {
        Pending cur = queue[h];  // copy: the queue grows below
        std::vector<Kid> rest;
        if (cur.kids.size() > 8) {  // keep 7, chain the rest behind a continuation node of the same level
            rest.assign(cur.kids.begin() + 7, cur.kids.end());
            cur.kids.resize(7);
            double m[4] = {0.0, 0.0, 0.0, 0.0};
            uint32_t count = 0;
            for (const Kid &k : rest) {
                const double w = (double)k.rec.cm.w;
                m[0] += w * (double)k.rec.cm.x;
                m[1] += w * (double)k.rec.cm.y;
                m[2] += w * (double)k.rec.cm.z;
                m[3] += w;
                count += k.rec.count;
            }
            NodeRec r;
            if (m[3] == 0.0) r.cm = make_float4(rest[0].rec.cm.x, rest[0].rec.cm.y, rest[0].rec.cm.z, 0.f);
            else r.cm = make_float4((float)(m[0] / m[3]), (float)(m[1] / m[3]), (float)(m[2] / m[3]), (float)m[3]);
            r.first_child = 0;
            r.nchild_level = (uint32_t)cur.level << 8;
            r.begin = rest[0].rec.begin;
            r.count = count;
            cur.kids.push_back({r, -2});
        }
        if (cur.kids.empty()) return fail(ctx, PCUDA_ERR_CUDA, "top tree: cell without children");
        const uint32_t first_child = (uint32_t)top.size();
        for (const Kid &k : cur.kids) {
            const uint32_t idx = (uint32_t)top.size();
            top.push_back(k.rec);
            if (k.cell >= 0) queue.push_back({idx, cells[k.cell].l, kids_of_cell(k.cell)});
            else if (k.cell == -2) queue.push_back({idx, cur.level, rest});
        }
        top[cur.me].first_child = top_base + first_child;
        top[cur.me].nchild_level = (uint32_t)cur.level << 8 | (uint32_t)cur.kids.size();
        if (top.size() > TOP_CAP) return fail(ctx, PCUDA_ERR_TREE_OVERFLOW, "top tree has %zu nodes", top.size());
    }
    roots.push_back(top_base);
    return PCUDA_OK;
}

// Steps 5-6 glue: boundary records -> host, merge, top tree + start nodes -> device.  `between`
// is enqueued after the boundary copy and overlaps the host merge.
template <class Between>
static int join_parts(pcuda_ctx *ctx, pcuda_forest *f, int parts, const uint32_t *node_base,
                      uint32_t top_base, Between between, ForestView *fv) {
    cudaStream_t st = ctx->stream;
    PartBases bases{};
    for (int q = 0; q < parts; ++q) bases.node_base[q] = node_base[q];
    collect_boundary<<<parts, 2 * TOP_LEVELS * 9, 0, st>>>(f->nodes.as<NodeRec>(), f->packs.as<PartPack>(),
                                                            bases, f->stage.as<BoundaryRec>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_stage, f->stage.p, (size_t)parts * TOP_LEVELS * 2 * sizeof(BoundaryRec),
                                        cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaEventRecord(f->ev_stage, st));
    PCUDA_TRY(between());
    PCUDA_CUDA_TRY(ctx, cudaEventSynchronize(f->ev_stage));
    std::vector<NodeRec> top;
    std::vector<uint32_t> roots;
    PCUDA_TRY(merge_top_tree(ctx, parts, f->h_packs, f->h_stage, node_base, top_base, top, roots));
    if (!top.empty())
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->nodes.as<NodeRec>() + top_base, top.data(), top.size() * sizeof(NodeRec),
                                            cudaMemcpyHostToDevice, st));
    if (!roots.empty())
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->roots.p, roots.data(), roots.size() * sizeof(uint32_t),
                                            cudaMemcpyHostToDevice, st));
    fv->nodes = f->nodes.as<NodeRec>();
    fv->src = f->sorted.as<float4>();
    fv->d_roots = f->roots.as<uint32_t>();
    fv->n_roots = (uint32_t)roots.size();
    return PCUDA_OK;
}

static int g_route = 0;  // accelerations to their owners: 0 = automatic, 1 = all-gather, 2 = all-to-all
// The all-to-all moves 16 B x N / world per rank instead of 12 B x N, but costs a synchronisation and
// three small launches more.  Measured on 8 B200s: N = 10M 6.01 ms against 5.85 ms per step with the
// all-gather, N = 80M 38.3 against 39.7 ms (2 GPUs, N = 10M: 0.2 ms slower) - hence only for large N.
static bool route_a2a(const pcuda_ctx *ctx, int world, size_t n_total) {
    return nccl_has_p2p(ctx) && world <= MAX_PARTS &&
           (g_route == 2 || (g_route == 0 && world >= 4 && n_total >= (size_t)32 << 20));
}

// d_idx: original index of each of this rank's n_rows traversal rows; cap: particles per owner block.
static int route_plan(pcuda_ctx *ctx, pcuda_forest *f, const uint32_t *d_idx, size_t n_rows, int world,
                      int rank, size_t cap, size_t n_own, RoutePlan *plan) {
    cudaStream_t st = ctx->stream;
    uint32_t *d_mat = f->route_cnt.as<uint32_t>();           // world rows of MAX_PARTS counts
    uint32_t *d_cursor = d_mat + MAX_PARTS * MAX_PARTS;      // MAX_PARTS
    uint32_t *d_row = d_mat + (size_t)rank * MAX_PARTS;
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_row, 0, MAX_PARTS * sizeof(uint32_t), st));
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_cursor, 0, MAX_PARTS * sizeof(uint32_t), st));
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * 8, (n_rows + 255) / 256));
    if (n_rows) owner_hist<<<grid, 256, 0, st>>>(d_idx, (int)n_rows, (uint32_t)cap, d_row);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, d_row, d_mat, MAX_PARTS * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_route, d_mat, (size_t)world * MAX_PARTS * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    OwnerOffsets so{};
    size_t s_off = 0, r_off = 0;
    for (int o = 0; o < world; ++o) {
        plan->send_off[o] = s_off;
        plan->send_cnt[o] = f->h_route[(size_t)rank * MAX_PARTS + o];
        so.off[o] = (uint32_t)s_off;
        s_off += plan->send_cnt[o];
        plan->recv_off[o] = r_off;
        plan->recv_cnt[o] = f->h_route[(size_t)o * MAX_PARTS + rank];
        r_off += plan->recv_cnt[o];
    }
    if (s_off != n_rows || r_off != n_own)
        return fail(ctx, PCUDA_ERR_NCCL, "routing plan is inconsistent (%zu of %zu rows out, %zu of %zu in)", s_off,
                    n_rows, r_off, n_own);
    plan->n_rows = n_rows;
    plan->n_recv = r_off;
    const size_t rows = std::max<size_t>(n_rows, 1), own = std::max<size_t>(n_own, 1);
    PCUDA_CUDA_TRY(ctx, f->route_pos.ensure(rows * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->route_idx_send.ensure(rows * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->route_acc_send.ensure(rows * 3 * sizeof(float)));
    PCUDA_CUDA_TRY(ctx, f->route_idx_recv.ensure(own * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->route_acc_recv.ensure(own * 3 * sizeof(float)));
    plan->d_pos = f->route_pos.as<uint32_t>();
    plan->d_acc_send = f->route_acc_send.as<float>();
    if (n_rows)
        owner_positions<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(d_idx, (int)n_rows, (uint32_t)cap, so, d_cursor,
                                                                         plan->d_pos, f->route_idx_send.as<uint32_t>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    return PCUDA_OK;
}

static int route_exchange(pcuda_ctx *ctx, pcuda_forest *f, const RoutePlan &plan, int world, size_t lo,
                          float *d_out) {
    size_t so[MAX_PARTS], sb[MAX_PARTS], ro[MAX_PARTS], rb[MAX_PARTS];
    for (int pass = 0; pass < 2; ++pass) {  // accelerations (12 B rows), then original indices (4 B)
        const size_t w = pass == 0 ? 12 : 4;
        for (int o = 0; o < world; ++o) {
            so[o] = plan.send_off[o] * w;
            sb[o] = plan.send_cnt[o] * w;
            ro[o] = plan.recv_off[o] * w;
            rb[o] = plan.recv_cnt[o] * w;
        }
        PCUDA_TRY(nccl_alltoallv(ctx, pass == 0 ? (const void *)f->route_acc_send.p : (const void *)f->route_idx_send.p, so,
                                 sb, pass == 0 ? f->route_acc_recv.p : f->route_idx_recv.p, ro, rb));
    }
    if (plan.n_recv) {
        scatter_rows<<<(unsigned)((plan.n_recv + 255) / 256), 256, 0, ctx->stream>>>(
            f->route_acc_recv.as<float>(), f->route_idx_recv.as<uint32_t>(), (int)plan.n_recv, (uint32_t)lo, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return PCUDA_OK;
}

// Diagnostic / test entry (one GPU): the parts are built one after the other ("virtual ranks"),
// joined and walked for all particles; out rows are in input order.  parts == 1 is the ordinary tree.
static int partitioned_dev(pcuda_ctx *ctx, const float *d_particles, size_t n, int parts, float theta,
                           float eps, float *d_out) {
    if (parts < 1 || parts > MAX_PARTS)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "parts must be in [1, %d]", MAX_PARTS);
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    if (ctx->order == 2)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "the partitioned build carries centre-of-mass nodes only");
    if (g_tpl != 2 || g_variant)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "the partitioned build is walked by traverse2_kernel only");
    if (n == 0) return PCUDA_OK;
    cudaStream_t st = ctx->stream;
    pcuda_forest *f = nullptr;
    PCUDA_TRY(forest_of(ctx, &f));
    uint32_t counts[MAX_PARTS] = {0};
    phase_begin(ctx, PH_BUILD);
    PCUDA_TRY(forest_partition(ctx, f, d_particles, n, parts, counts));
    size_t slot = 1, total = 0;
    for (int q = 0; q < parts; ++q) {
        slot = std::max<size_t>(slot, counts[q]);
        total += counts[q];
    }
    if (total != n) return fail(ctx, PCUDA_ERR_CUDA, "partition lost particles (%zu of %zu)", total, n);
    PCUDA_CUDA_TRY(ctx, f->sorted.ensure((size_t)parts * slot * sizeof(float4)));
    PCUDA_CUDA_TRY(ctx, f->perm.ensure((size_t)parts * slot * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->keys.ensure((size_t)parts * slot * sizeof(uint64_t)));
    uint32_t node_base[MAX_PARTS] = {0};
    size_t next = 0;
    for (int q = 0; q < parts; ++q) {
        PCUDA_TRY(forest_build_part(ctx, f, d_particles, n, q, counts[q], slot, f->packs.as<PartPack>() + q));
        node_base[q] = (uint32_t)next;
        if (counts[q] == 0) continue;
        const pcuda_tree *t = f->local;
        const size_t need = (next + t->n_nodes + TOP_CAP) * sizeof(NodeRec);
        if (need > f->nodes.cap) {  // grow, keeping the parts already placed
            DevBuf bigger;
            PCUDA_CUDA_TRY(ctx, bigger.ensure(std::max(need, ((size_t)parts * t->n_nodes + TOP_CAP) * sizeof(NodeRec))));
            cudaError_t e = next ? cudaMemcpyAsync(bigger.p, f->nodes.p, next * sizeof(NodeRec),
                                                   cudaMemcpyDeviceToDevice, st)
                                 : cudaSuccess;
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) bigger.release();
            PCUDA_CUDA_TRY(ctx, e);
            f->nodes.release();
            f->nodes = bigger;
        }
        copy_rebase_nodes<<<(unsigned)((t->n_nodes + 255) / 256), 256, 0, st>>>(
            t->nodes.as<NodeRec>(), (uint32_t)t->n_nodes, (uint32_t)next, (uint32_t)(q * slot),
            f->nodes.as<NodeRec>() + next);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->sorted.as<float4>() + q * slot, t->sorted.p,
                                            counts[q] * sizeof(float4), cudaMemcpyDeviceToDevice, st));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->perm.as<uint32_t>() + q * slot, t->d_perm(),
                                            counts[q] * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->keys.as<uint64_t>() + q * slot, t->d_keys(),
                                            counts[q] * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        next += t->n_nodes;
    }
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_packs, f->packs.p, parts * sizeof(PartPack), cudaMemcpyDeviceToHost, st));
    ForestView fv{};
    PCUDA_TRY(join_parts(ctx, f, parts, node_base, (uint32_t)next, [] { return (int)PCUDA_OK; }, &fv));
    phase_end(ctx, PH_BUILD);
    phase_begin(ctx, PH_COMPUTE);
    for (int q = 0; q < parts; ++q) {
        if (counts[q] == 0) continue;
        PCUDA_TRY(traverse_sorted(ctx, f->local, f->sorted.as<float4>() + q * slot,
                                  f->keys.as<uint64_t>() + q * slot, f->perm.as<uint32_t>() + q * slot,
                                  counts[q], theta, eps, d_out, nullptr, &fv));
    }
    phase_end(ctx, PH_COMPUTE);
    return PCUDA_OK;
}

static int g_forest = 0;  // multi-GPU Barnes-Hut build: 0 = as the context says, 1 = partitioned, 2 = replicated

// Multi-GPU step with the partitioned build: d_gathered already holds all n_total records.  Rank r
// builds the tree of the r-th key range, the trees are exchanged and joined, and rank r walks the
// result for the targets of its own range.  Same result routing as the replicated path.
static int sharded_forest_dev(pcuda_ctx *ctx, int world, int rank, size_t n_total, size_t lo, size_t hi,
                              float theta, float eps, const float *d_gathered, float *d_out) {
    cudaStream_t st = ctx->stream;
    pcuda_forest *f = nullptr;
    PCUDA_TRY(forest_of(ctx, &f));
    uint32_t counts[MAX_PARTS] = {0};
    phase_begin(ctx, PH_BUILD);
    PCUDA_TRY(forest_partition(ctx, f, d_gathered, n_total, world, counts));
    size_t slot = 1, total = 0;
    for (int q = 0; q < world; ++q) {
        slot = std::max<size_t>(slot, counts[q]);
        total += counts[q];
    }
    if (total != n_total)
        return fail(ctx, PCUDA_ERR_CUDA, "partition lost particles (%zu of %zu)", total, n_total);
    const size_t mine = counts[rank];
    PartPack *d_packs = f->packs.as<PartPack>();
    PCUDA_TRY(forest_build_part(ctx, f, d_gathered, n_total, rank, mine, slot, d_packs + rank));
    phase_end(ctx, PH_BUILD);
    const pcuda_tree *t = f->local;

    // exchange: packs (node counts, level tables, boundary moments) -> common slot size; node
    // records and permutations into equal slots
    phase_begin(ctx, PH_COMM3);
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, d_packs + rank, d_packs, sizeof(PartPack)));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_packs, d_packs, world * sizeof(PartPack), cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    size_t node_slot = 1;
    for (int q = 0; q < world; ++q) node_slot = std::max<size_t>(node_slot, f->h_packs[q].n_nodes);
    if ((size_t)world * node_slot + TOP_CAP > 0xfffffff0ull || (size_t)world * slot > 0xfffffff0ull)
        return fail(ctx, PCUDA_ERR_TREE_OVERFLOW, "joined tree does not fit 32-bit indices");
    const uint32_t my_nodes = (uint32_t)t->n_nodes;
    if (f->h_packs[rank].n_nodes != my_nodes) return fail(ctx, PCUDA_ERR_CUDA, "pack exchange is inconsistent");
    PCUDA_CUDA_TRY(ctx, f->nodes.ensure(((size_t)world * node_slot + TOP_CAP) * sizeof(NodeRec)));
    PCUDA_CUDA_TRY(ctx, f->perm.ensure((size_t)world * slot * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->sorted.ensure((size_t)world * slot * sizeof(float4)));
    PCUDA_CUDA_TRY(ctx, f->acc.ensure((size_t)world * slot * 3 * sizeof(float)));
    NodeRec *my_node_slot = f->nodes.as<NodeRec>() + (size_t)rank * node_slot;
    uint32_t *my_perm_slot = f->perm.as<uint32_t>() + (size_t)rank * slot;
    if (my_nodes)
        copy_rebase_nodes<<<(my_nodes + 255) / 256, 256, 0, st>>>(
            t->nodes.as<NodeRec>(), my_nodes, (uint32_t)(rank * node_slot), (uint32_t)(rank * slot),
            my_node_slot);
    copy_pad_perm<<<(unsigned)((slot + 255) / 256), 256, 0, st>>>(
        mine ? t->d_perm() : nullptr, (uint32_t)mine, (uint32_t)slot, my_perm_slot);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, my_node_slot, f->nodes.p, node_slot * sizeof(NodeRec)));
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, my_perm_slot, f->perm.p, slot * sizeof(uint32_t)));
    const size_t n_slots = (size_t)world * slot;
    uint32_t node_base[MAX_PARTS] = {0};
    for (int q = 0; q < world; ++q) node_base[q] = (uint32_t)(q * node_slot);
    ForestView fv{};
    PCUDA_TRY(join_parts(
        ctx, f, world, node_base, (uint32_t)(world * node_slot),
        [&]() -> int {  // overlaps the host merge
            gather_forest<<<(unsigned)((n_slots + 255) / 256), 256, 0, st>>>(
                reinterpret_cast<const float4 *>(d_gathered), f->perm.as<uint32_t>(), n_slots,
                f->sorted.as<float4>());
            PCUDA_CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
            return PCUDA_OK;
        },
        &fv));

    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    if (route_a2a(ctx, world, n_total)) {  // every row goes to its owner only
        RoutePlan plan;
        PCUDA_TRY(route_plan(ctx, f, mine ? t->d_perm() : nullptr, mine, world, rank, cap, hi - lo, &plan));
        phase_end(ctx, PH_COMM3);
        phase_begin(ctx, PH_COMPUTE);
        if (mine)
            PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>(), t->d_keys(), plan.d_pos, mine, theta, eps,
                                      plan.d_acc_send, nullptr, &fv));
        phase_end(ctx, PH_COMPUTE);
        phase_begin(ctx, PH_COMM2);
        PCUDA_TRY(route_exchange(ctx, f, plan, world, lo, d_out));
        phase_end(ctx, PH_COMM2);
        return PCUDA_OK;
    }
    phase_end(ctx, PH_COMM3);
    float *acc = f->acc.as<float>();
    phase_begin(ctx, PH_COMPUTE);
    if (mine)
        PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>(), t->d_keys(), nullptr, mine, theta, eps,
                                  acc + (size_t)rank * slot * 3, nullptr, &fv));
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_COMM2);
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, acc + (size_t)rank * slot * 3, acc, slot * 12));
    if (hi > lo) {
        pick_owned_rows<<<(unsigned)((n_slots + 255) / 256), 256, 0, st>>>(
            acc, f->perm.as<uint32_t>(), (int)n_slots, (uint32_t)lo, (uint32_t)hi, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    phase_end(ctx, PH_COMM2);
    return PCUDA_OK;
}

static int sharded_dev(pcuda_ctx *ctx, const float *d_local, size_t n_local, size_t n_total,
                       float theta, float eps, float *d_gathered, float *d_out) {
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    const size_t lo = std::min(n_total, (size_t)rank * cap), hi = std::min(n_total, lo + cap);
    if (n_local != hi - lo)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "rank %d of %d must own %zu of %zu particles (contiguous blocks of %zu), got %zu",
                    rank, world, hi - lo, n_total, cap, n_local);
    float *slot = d_gathered + (size_t)rank * cap * 4;
    phase_begin(ctx, PH_COMM);
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(slot, d_local, n_local * 16, cudaMemcpyDeviceToDevice,
                                            ctx->stream));
    if (world > 1) PCUDA_TRY(pcuda_comm_allgather_dev(ctx, slot, d_gathered, cap * 16));
    phase_end(ctx, PH_COMM);
    const int how = g_forest ? g_forest : ctx->bh_build;  // 0 = automatic: partitioned from 4 GPUs on
    const bool forest = how == 1 || (how == 0 && world >= 4);
    if (forest && world > 1 && world <= MAX_PARTS && n_total >= (size_t)world && ctx->order == 1 &&
        g_tpl == 2 && !g_variant)
        return sharded_forest_dev(ctx, world, rank, n_total, lo, hi, theta, eps, d_gathered, d_out);
    if (!ctx->call_tree) ctx->call_tree = new pcuda_tree();
    pcuda_tree *t = ctx->call_tree;
    phase_begin(ctx, PH_BUILD);
    PCUDA_TRY(build_dim(ctx, t, 3, d_gathered, n_total));
    phase_end(ctx, PH_BUILD);
    if (world == 1) {
        phase_begin(ctx, PH_COMPUTE);
        PCUDA_TRY(traverse(ctx, t, nullptr, n_total, theta, eps, d_out));
        phase_end(ctx, PH_COMPUTE);
        return PCUDA_OK;
    }
    if (route_a2a(ctx, world, n_total)) {  // every row goes to its owner only
        pcuda_forest *f = nullptr;
        PCUDA_TRY(forest_of(ctx, &f));
        RoutePlan plan;
        phase_begin(ctx, PH_COMM3);
        PCUDA_TRY(route_plan(ctx, f, t->d_perm() + lo, n_local, world, rank, cap, n_local, &plan));
        phase_end(ctx, PH_COMM3);
        phase_begin(ctx, PH_COMPUTE);
        if (n_local)
            PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>() + lo, t->d_keys() + lo, plan.d_pos, n_local,
                                      theta, eps, plan.d_acc_send));
        phase_end(ctx, PH_COMPUTE);
        phase_begin(ctx, PH_COMM2);
        PCUDA_TRY(route_exchange(ctx, f, plan, world, lo, d_out));
        phase_end(ctx, PH_COMM2);
        return PCUDA_OK;
    }
    // accelerations of all particles in key order, world * cap rows; this rank fills rows [lo, hi)
    PCUDA_CUDA_TRY(ctx, ctx->d_misc.ensure((size_t)world * cap * 3 * sizeof(float)));
    float *acc_sorted = ctx->d_misc.as<float>();
    phase_begin(ctx, PH_COMPUTE);
    if (n_local)
        PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>() + lo, t->d_keys() + lo, nullptr, n_local,
                                  theta, eps, acc_sorted + lo * 3));
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_COMM2);  // the second exchange of the call; reported inside comm_ms
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, acc_sorted + (size_t)rank * cap * 3, acc_sorted, cap * 12));
    if (n_local) {
        pick_owned_rows<<<(unsigned)((n_total + 255) / 256), 256, 0, ctx->stream>>>(
            acc_sorted, t->d_perm(), (int)n_total, (uint32_t)lo, (uint32_t)hi, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    phase_end(ctx, PH_COMM2);
    return PCUDA_OK;
}

static int oneshot_host(pcuda_ctx *ctx, uint32_t dim, const float *aff, size_t na, const float *src,
                        size_t nb, float theta, float eps, float *out) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if ((na && !out) || (nb && !src))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (!aff && na != nb)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "affected == NULL means affected == affecting, but n_affected != n_affecting");
    if (na > 0x7fffffffull || nb > 0x7fffffffull)  // before any buffer is touched
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (na == 0) return PCUDA_OK;
    const size_t src_bytes = nb * (dim + 1) * sizeof(float), tgt_bytes = na * dim * sizeof(float);
    phase_begin(ctx, PH_UPLOAD);
    float *d_src = nullptr, *d_tgt = nullptr;
    if (nb) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(src_bytes));
        d_src = ctx->d_affecting.as<float>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (aff) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affected.ensure(tgt_bytes));
        d_tgt = ctx->d_affected.as<float>();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(d_tgt, aff, tgt_bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(tgt_bytes));
    phase_end(ctx, PH_UPLOAD);
    PCUDA_TRY(oneshot_dev(ctx, dim, d_tgt, na, d_src, nb, theta, eps, ctx->d_out.as<float>()));
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out.p, tgt_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    PCUDA_TRY(timings_collect(ctx));
    return read_counters(ctx);
}

}  // namespace bh

int bh_enqueue_f32(pcuda_ctx *ctx, int dim, const float *d_tgt, int tgt_stride, size_t na,
                   const float *d_src, size_t nb, float theta, float softening, float *d_out) {
    return bh::oneshot_dev(ctx, (uint32_t)dim, tgt_stride ? d_tgt : nullptr, na, d_src, nb, theta,
                           softening, d_out, tgt_stride);
}

int bh_enqueue_f64(pcuda_ctx *ctx, int dim, const double *d_tgt, int tgt_stride, size_t na,
                   const double *d_src, size_t nb, double theta, double softening, double *d_out) {
    return bh::oneshot_dev64(ctx, (uint32_t)dim, tgt_stride ? d_tgt : nullptr, na, d_src, nb, theta,
                             softening, d_out, tgt_stride);
}

int bh_debug_set(const char *key, int value) {
    const std::string k = key ? key : "";
    if (k == "bh_seg_max" && value >= 32 && value <= bh::SEG_MAX_LIMIT && value % 32 == 0) {
        bh::g_seg_max = value;
        return PCUDA_OK;
    }
    if (k == "bh_tpl" && (value == 1 || value == 2)) {
        bh::g_tpl = value;
        return PCUDA_OK;
    }
    if (k == "bh_small_level" && value >= 0) {
        bh::g_small_level = (uint32_t)value;
        return PCUDA_OK;
    }
    if (k == "bh_level_build" && value >= 0 && value <= 2) {
        bh::g_level_build = value;
        return PCUDA_OK;
    }
    if (k == "bh_variant" && value >= 0 && value <= 3) {
        bh::g_variant = value;
        return PCUDA_OK;
    }
    if (k == "bh_count") {
        bh::g_count = value != 0;
        return PCUDA_OK;
    }
    if (k == "bh_forest" && value >= 0 && value <= 2) {
        bh::g_forest = value;
        return PCUDA_OK;
    }
    if (k == "bh_route" && value >= 0 && value <= 2) {
        bh::g_route = value;
        return PCUDA_OK;
    }
    return PCUDA_ERR_INVALID_ARGUMENT;
}

void tree_free(pcuda_ctx *ctx, pcuda_tree *t) {
    if (!t) return;
    (void)ctx;
    DevBuf *bufs[] = {&t->keys[0], &t->keys[1], &t->perm[0], &t->perm[1], &t->sorted, &t->nodes,
                      &t->moments, &t->d_frame, &t->scan_in, &t->scan_out, &t->cub_tmp, &t->partial,
                      &t->sorted64, &t->quad64, &t->quad, &t->rb};
    for (DevBuf *b : bufs) b->release();
    delete t;
}

}  // namespace pcuda

using namespace pcuda;

extern "C" {

int pcuda_barneshut_f32x3(pcuda_ctx *ctx, const float *aff, size_t na, const float *src, size_t nb,
                          float theta, float softening, int checked, float *out) {
    (void)checked;  // a zero-distance pair never contributes in Barnes-Hut (sequential.rs:485-487)
    return bh::oneshot_host(ctx, 3, aff, na, src, nb, theta, softening, out);
}

int pcuda_barneshut_f32x2(pcuda_ctx *ctx, const float *aff, size_t na, const float *src, size_t nb,
                          float theta, float softening, int checked, float *out) {
    (void)checked;
    return bh::oneshot_host(ctx, 2, aff, na, src, nb, theta, softening, out);
}

int pcuda_barneshut_f64x3(pcuda_ctx *ctx, const double *aff, size_t na, const double *src, size_t nb,
                          double theta, double softening, int checked, double *out) {
    (void)checked;  // a pair at zero distance contributes nothing either way (sequential.rs:485-487)
    return bh::oneshot_host64(ctx, 3, aff, na, src, nb, theta, softening, out);
}

int pcuda_barneshut_f64x2(pcuda_ctx *ctx, const double *aff, size_t na, const double *src, size_t nb,
                          double theta, double softening, int checked, double *out) {
    (void)checked;
    return bh::oneshot_host64(ctx, 2, aff, na, src, nb, theta, softening, out);
}

static int bh_dev64(pcuda_ctx *ctx, uint32_t dim, const double *d_aff, size_t na, const double *d_src,
                    size_t nb, double theta, double softening, double *d_out) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (na == 0) return PCUDA_OK;
    int s = bh::oneshot_dev64(ctx, dim, d_aff, na, d_src, nb, theta, softening, d_out);
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_barneshut_f64x3_dev(pcuda_ctx *ctx, const double *d_aff, size_t na, const double *d_src,
                              size_t nb, double theta, double softening, int checked, double *d_out) {
    (void)checked;
    return bh_dev64(ctx, 3, d_aff, na, d_src, nb, theta, softening, d_out);
}

int pcuda_barneshut_f64x2_dev(pcuda_ctx *ctx, const double *d_aff, size_t na, const double *d_src,
                              size_t nb, double theta, double softening, int checked, double *d_out) {
    (void)checked;
    return bh_dev64(ctx, 2, d_aff, na, d_src, nb, theta, softening, d_out);
}

static int bh_dev(pcuda_ctx *ctx, uint32_t dim, const float *d_aff, size_t na, const float *d_src,
                  size_t nb, float theta, float softening, float *d_out) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int s = bh::oneshot_dev(ctx, dim, d_aff, na, d_src, nb, theta, softening, d_out);
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_barneshut_f32x3_dev(pcuda_ctx *ctx, const float *d_aff, size_t na, const float *d_src,
                              size_t nb, float theta, float softening, int checked, float *d_out) {
    (void)checked;
    return bh_dev(ctx, 3, d_aff, na, d_src, nb, theta, softening, d_out);
}

int pcuda_barneshut_f32x2_dev(pcuda_ctx *ctx, const float *d_aff, size_t na, const float *d_src,
                              size_t nb, float theta, float softening, int checked, float *d_out) {
    (void)checked;
    return bh_dev(ctx, 2, d_aff, na, d_src, nb, theta, softening, d_out);
}

int pcuda_barneshut_f32x3_sharded_dev(pcuda_ctx *ctx, const float *d_local_xyzm, size_t n_local,
                                      size_t n_total, float theta, float softening, int checked,
                                      float *d_gathered_xyzm, float *d_out_xyz) {
    (void)checked;
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int s = bh::sharded_dev(ctx, d_local_xyzm, n_local, n_total, theta, softening, d_gathered_xyzm,
                            d_out_xyz);
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_barneshut_f32x3_sharded(pcuda_ctx *ctx, const float *local_xyzm, size_t n_local,
                                  size_t n_total, float theta, float softening, int checked,
                                  float *out_xyz) {
    (void)checked;
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n_local && (!local_xyzm || !out_xyz))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(cap * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_packed_src.ensure((size_t)world * cap * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(cap * 12));
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affecting.p, local_xyzm, n_local * 16,
                                            cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    PCUDA_TRY(bh::sharded_dev(ctx, ctx->d_affecting.as<float>(), n_local, n_total, theta, softening,
                              ctx->d_packed_src.as<float>(), ctx->d_out.as<float>()));
    phase_begin(ctx, PH_DOWNLOAD);
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out_xyz, ctx->d_out.p, n_local * 12, cudaMemcpyDeviceToHost,
                                            ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    PCUDA_TRY(timings_collect(ctx));
    return bh::read_counters(ctx);
}

int pcuda_barneshut_f32x3_partitioned_dev(pcuda_ctx *ctx, const float *d_xyzm, size_t n, int parts,
                                          float theta, float softening, int checked, float *d_out_xyz) {
    (void)checked;
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n && (!d_xyzm || !d_out_xyz))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    int s = bh::partitioned_dev(ctx, d_xyzm, n, parts, theta, softening, d_out_xyz);
    ctx->timings.kernel_launches = ctx->launches;
    return s;
}

int pcuda_barneshut_f32x3_partitioned(pcuda_ctx *ctx, const float *xyzm, size_t n, int parts, float theta,
                                      float softening, int checked, float *out_xyz) {
    (void)checked;
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n && (!xyzm || !out_xyz))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (n == 0) return PCUDA_OK;
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(n * 16));
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(n * 12));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affecting.p, xyzm, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    PCUDA_TRY(bh::partitioned_dev(ctx, ctx->d_affecting.as<float>(), n, parts, theta, softening,
                                  ctx->d_out.as<float>()));
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out_xyz, ctx->d_out.p, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    return timings_collect(ctx);
}

// Test hook (not in the stable header): the host-side merge of the partitioned build on caller-made
// inputs, no device involved.  packs: parts x pcuda::bh::PartPack; stage: parts x 22 x 2 x BoundaryRec;
// top_out: room for top_cap 32-byte node records; roots_out: room for roots_cap indices.
int pcuda_debug_merge_top_tree(int parts, const void *packs, const void *stage, const uint32_t *node_base,
                               uint32_t top_base, void *top_out, uint32_t top_cap, uint32_t *n_top,
                               uint32_t *roots_out, uint32_t roots_cap, uint32_t *n_roots) {
    if (parts < 1 || parts > bh::MAX_PARTS || !packs || !stage || !node_base || !n_top || !n_roots)
        return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "bad arguments");
    std::vector<bh::NodeRec> top;
    std::vector<uint32_t> roots;
    PCUDA_TRY(bh::merge_top_tree(nullptr, parts, static_cast<const bh::PartPack *>(packs),
                                 static_cast<const bh::BoundaryRec *>(stage), node_base, top_base, top, roots));
    if (top.size() > top_cap || roots.size() > roots_cap)
        return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "output buffers too small");
    if (!top.empty()) memcpy(top_out, top.data(), top.size() * sizeof(bh::NodeRec));
    if (!roots.empty()) memcpy(roots_out, roots.data(), roots.size() * sizeof(uint32_t));
    *n_top = (uint32_t)top.size();
    *n_roots = (uint32_t)roots.size();
    return PCUDA_OK;
}

int pcuda_tree_build_f32(pcuda_ctx *ctx, uint32_t dim, const float *affecting, size_t n,
                         pcuda_tree **out) {
    if (!ctx || !out) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    if (dim != 2 && dim != 3) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    if (n && !affecting) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    const size_t bytes = n * (dim + 1) * sizeof(float);
    phase_begin(ctx, PH_UPLOAD);
    if (n) {
        PCUDA_CUDA_TRY(ctx, ctx->d_affecting.ensure(bytes));
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affecting.p, affecting, bytes, cudaMemcpyHostToDevice,
                                            ctx->stream));
    }
    phase_end(ctx, PH_UPLOAD);
    pcuda_tree *t = new pcuda_tree();
    phase_begin(ctx, PH_BUILD);
    int s = bh::build_dim(ctx, t, dim, ctx->d_affecting.as<float>(), n);
    if (s != PCUDA_OK) {
        tree_free(ctx, t);
        return s;
    }
    phase_end(ctx, PH_BUILD);
    s = timings_collect(ctx);
    if (s != PCUDA_OK) {
        tree_free(ctx, t);
        return s;
    }
    *out = t;
    return PCUDA_OK;
}

// Morton keys + stable sort permutation alone (SURVEY.md 8b `pcuda_morton_*`): the first half of
// the tree build (root cube per BoundingBox::square_with, tree/partition.rs:136-153; quantisation;
// stable radix sort), read back to the host.  keys_out[i] = i-th smallest key, perm_out[i] = index
// of the particle that holds it (ties in input order).
static int morton_host(pcuda_ctx *ctx, uint32_t dim, const float *particles, size_t n, uint64_t *keys_out,
                       uint32_t *perm_out, pcuda_tree_info *frame_out) {
    if (!ctx) return fail(nullptr, PCUDA_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n && (!particles || !keys_out || !perm_out))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    if (n > 0x7fffffffull) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "particle count exceeds 2^31-1");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    pcuda_tree t;
    int s = PCUDA_OK;
    do {
        if (n == 0) break;
        const size_t bytes = n * (dim + 1) * sizeof(float);
        phase_begin(ctx, PH_UPLOAD);
        if (ctx->d_affecting.ensure(bytes) != cudaSuccess) {
            s = fail(ctx, PCUDA_ERR_OUT_OF_MEMORY, "out of device memory");
            break;
        }
        cudaMemcpyAsync(ctx->d_affecting.p, particles, bytes, cudaMemcpyHostToDevice, ctx->stream);
        phase_end(ctx, PH_UPLOAD);
        phase_begin(ctx, PH_BUILD);
        s = bh::build_dim(ctx, &t, dim, ctx->d_affecting.as<float>(), n, true);
        if (s != PCUDA_OK) break;
        phase_end(ctx, PH_BUILD);
        phase_begin(ctx, PH_DOWNLOAD);
        cudaMemcpyAsync(keys_out, t.d_keys(), n * 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(perm_out, t.d_perm(), n * 4, cudaMemcpyDeviceToHost, ctx->stream);
        phase_end(ctx, PH_DOWNLOAD);
        s = timings_collect(ctx);
        if (s == PCUDA_OK && cudaGetLastError() != cudaSuccess) s = fail(ctx, PCUDA_ERR_CUDA, "copy failed");
    } while (0);
    if (s == PCUDA_OK && frame_out) {
        t.dim = (int)dim;
        t.bits = dim == 3 ? 21 : 31;
        t.n = n;
        pcuda_tree_info_get(&t, frame_out);
    }
    cudaStreamSynchronize(ctx->stream);
    for (pcuda::DevBuf *b : {&t.keys[0], &t.keys[1], &t.perm[0], &t.perm[1], &t.sorted, &t.nodes, &t.moments,
                             &t.d_frame, &t.scan_in, &t.scan_out, &t.cub_tmp, &t.partial})
        b->release();
    return s;
}

int pcuda_morton_f32x3(pcuda_ctx *ctx, const float *particles_xyzm, size_t n, uint64_t *keys_out,
                       uint32_t *perm_out, pcuda_tree_info *frame_out) {
    return morton_host(ctx, 3, particles_xyzm, n, keys_out, perm_out, frame_out);
}

int pcuda_morton_f32x2(pcuda_ctx *ctx, const float *particles_xym, size_t n, uint64_t *keys_out,
                       uint32_t *perm_out, pcuda_tree_info *frame_out) {
    return morton_host(ctx, 2, particles_xym, n, keys_out, perm_out, frame_out);
}

int pcuda_tree_info_get(const pcuda_tree *t, pcuda_tree_info *out) {
    if (!t || !out) return PCUDA_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof *out);
    out->n_particles = t->n;
    out->n_nodes = t->n_nodes;
    out->n_levels = (uint32_t)t->n_levels;
    out->leaf_size = t->leaf_size;
    out->dim = (uint32_t)t->dim;
    out->bits = (uint32_t)t->bits;
    for (int k = 0; k < 3; ++k) out->origin[k] = t->frame.origin[k];
    out->extent = t->frame.ext;
    out->inv = t->frame.inv;
    return PCUDA_OK;
}

int pcuda_tree_read(pcuda_ctx *ctx, const pcuda_tree *t, int which, void *dst, size_t dst_bytes) {
    if (!ctx || !t || (!dst && dst_bytes))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(ctx->device);
    const size_t n = t->n, m = t->n_nodes;
    size_t need = 0;
    switch (which) {
        case PCUDA_TREE_KEYS: need = n * 8; break;
        case PCUDA_TREE_PERM: need = n * 4; break;
        case PCUDA_TREE_NODE_BEGIN:
        case PCUDA_TREE_NODE_COUNT:
        case PCUDA_TREE_NODE_LEVEL:
        case PCUDA_TREE_NODE_FIRST_CHILD:
        case PCUDA_TREE_NODE_NUM_CHILDREN: need = m * 4; break;
        case PCUDA_TREE_NODE_COM_MASS: need = m * (t->dim + 1) * 4; break;
        default: return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "unknown tree array %d", which);
    }
    if (dst_bytes < need)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "destination holds %zu bytes, %zu needed", dst_bytes, need);
    if (need == 0) return PCUDA_OK;
    if (which == PCUDA_TREE_KEYS || which == PCUDA_TREE_PERM) {
        const void *src = which == PCUDA_TREE_KEYS ? (const void *)t->d_keys() : (const void *)t->d_perm();
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return PCUDA_OK;
    }
    std::vector<bh::NodeRec> h(m);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), t->nodes.p, m * sizeof(bh::NodeRec), cudaMemcpyDeviceToHost,
                                        ctx->stream));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t *u = static_cast<uint32_t *>(dst);
    float *f = static_cast<float *>(dst);
    for (size_t j = 0; j < m; ++j) {
        const bh::NodeRec &r = h[j];
        switch (which) {
            case PCUDA_TREE_NODE_BEGIN: u[j] = r.begin; break;
            case PCUDA_TREE_NODE_COUNT: u[j] = r.count; break;
            case PCUDA_TREE_NODE_LEVEL: u[j] = r.nchild_level >> 8; break;
            case PCUDA_TREE_NODE_FIRST_CHILD: u[j] = r.first_child; break;
            case PCUDA_TREE_NODE_NUM_CHILDREN: u[j] = r.nchild_level & 0xffu; break;
            default:
                if (t->dim == 3) {
                    f[4 * j + 0] = r.cm.x; f[4 * j + 1] = r.cm.y; f[4 * j + 2] = r.cm.z; f[4 * j + 3] = r.cm.w;
                } else {
                    f[3 * j + 0] = r.cm.x; f[3 * j + 1] = r.cm.y; f[3 * j + 2] = r.cm.w;
                }
        }
    }
    return PCUDA_OK;
}

int pcuda_tree_traverse_f32(pcuda_ctx *ctx, const pcuda_tree *t, const float *affected, size_t na,
                            float theta, float softening, int checked, float *out) {
    (void)checked;
    if (!ctx || !t) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    if (na && (!affected || !out)) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL buffer with non-zero count");
    DeviceGuard guard(ctx->device);
    timings_reset(ctx);
    if (na == 0) return PCUDA_OK;
    const size_t bytes = na * t->dim * sizeof(float);
    phase_begin(ctx, PH_UPLOAD);
    PCUDA_CUDA_TRY(ctx, ctx->d_affected.ensure(bytes));
    PCUDA_CUDA_TRY(ctx, ctx->d_out.ensure(bytes));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_affected.p, affected, bytes, cudaMemcpyHostToDevice, ctx->stream));
    phase_end(ctx, PH_UPLOAD);
    phase_begin(ctx, PH_COMPUTE);
    PCUDA_TRY(bh::traverse(ctx, t, ctx->d_affected.as<float>(), na, theta, softening, ctx->d_out.as<float>()));
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_DOWNLOAD);
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    phase_end(ctx, PH_DOWNLOAD);
    PCUDA_TRY(timings_collect(ctx));
    return bh::read_counters(ctx);
}

int pcuda_tree_last_counters(pcuda_ctx *ctx, uint64_t counters[5]) {
    if (!ctx || !counters) return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "NULL argument");
    DeviceGuard guard(ctx->device);
    PCUDA_TRY(bh::read_counters(ctx));
    for (int i = 0; i < 5; ++i) counters[i] = ctx->last_counters[i];
    return PCUDA_OK;
}

void pcuda_tree_destroy(pcuda_ctx *ctx, pcuda_tree *t) {
    if (!t) return;
    if (ctx) {
        DeviceGuard guard(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        tree_free(ctx, t);
    } else {
        tree_free(nullptr, t);
    }
}

}  // extern "C"
