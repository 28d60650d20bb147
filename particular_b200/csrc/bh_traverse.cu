// bh_traverse.cu — K5 of the Barnes-Hut path: target groups and the warp-cooperative theta-traversals
// (f32 one / two targets per lane, quadrupole nodes, double precision); host-side traversal drivers.
// See barneshut.cu for the overview of the path and the reference lines it replaces
// (particular/src/sequential.rs:466-505).
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>

#include "bh.cuh"
#include "ptx.cuh"

namespace pcuda {
namespace bh {

// ------------------------------------------------------------------------------------------------
// K5a: target groups.  The targets are walked in key order in groups of at most 32 that never
// straddle a coarse cell boundary: a SEGMENT is a maximal cell (key prefix) holding at most
// `seg_max` targets (found from the keys alone: adjacent keys first differ at digit L[i], and the
// cell they share is counted by scanning L to both sides), and every segment is cut into equal
// chunks of <= 32 consecutive targets.  Without this, 32 consecutive keys that cross e.g. the
// centre of a Plummer sphere have a bounding box spanning the core and open millions of nodes.
constexpr int GROUP_BLOCK = 256;

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
static_assert(MAX_ROOTS == STACK_CAP - STACK_RESERVE - 32, "start nodes of a forest walk");

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
    const uint32_t *n_roots_dev;  // forest walks: the number of start nodes when only the device knows it
    int accumulate;               // forest walks: add to the output rows (second phase of a two-phase walk)
    uint32_t reserve_from, reserve_cap;  // forest walks: blocks on SMs >= reserve_from exit (at most reserve_cap)
    const uint32_t *chain;               // forest walks: level table of the tree whose first / last nodes are shares
    float share_scale;                   // forest walks: 1 + sqrt(3) theta
    const uint32_t *stop;                // forest walks: no more groups are taken once this word is non-zero
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

    if (FOREST && a.reserve_cap) {
        // Two-phase walk, first phase: the grid fills the GPU, and the blocks that land on the last SMs
        // leave at once, so that the kernels of the other stream (tree pruning, NCCL) find free SMs.  The
        // groups are handed out dynamically, the remaining blocks do all the work; at most reserve_cap
        // blocks leave, wherever the scheduler puts them.
        __shared__ int s_quit;
        if (threadIdx.x == 0)
            s_quit = ptx::smid() >= a.reserve_from && atomicAdd(a.counters + 6, 1ull) < (unsigned long long)a.reserve_cap;
        __syncthreads();
        if (s_quit) return;
    }
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
        if (FOREST) {  // (a.stop: the first phase of a two-phase walk ends when the second can begin)
            if (lane == 0) g = a.stop && *reinterpret_cast<const volatile uint32_t *>(a.stop) ? 0xffffffffu : atomicAdd(a.work, 1u);
        } else {
            if (lane == 0) g = atomicAdd(a.work, 1u);
        }
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
            sp = a.n_roots_dev ? (int)*a.n_roots_dev : (int)a.n_roots;
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
                const int level = FOREST ? (int)(nd.nchild_level >> 8 & 0xffu) : (int)(nd.nchild_level >> 8);
                float w = ext * __int_as_float((127 - level) << 23);
                if (FOREST) {
                    // Two-phase walk: a node that may be one side's share of a cell (the others' share: bit 31;
                    // this rank's own: the first / last node of a level of its tree, a.chain = the level
                    // table) is accepted only where the whole cell would be.  The cell's centre of mass lies
                    // in the same cube, at most sqrt(3) w from the share's, and the distance to a box is
                    // 1-Lipschitz: theta (d - sqrt(3) w) >= w, i.e. the test with w (1 + sqrt(3) theta).
                    bool share = (nd.nchild_level & NODE_SHARE) != 0u;
                    if (a.chain) share |= id == __ldg(a.chain + level) || id + 1u == __ldg(a.chain + level + 1);
                    if (share) w *= a.share_scale;
                }
                open = a.theta2 * d2 < w * w;
            }
            const uint32_t nc = nd.nchild_level & 0xffu;
            const bool open_internal = has && open && nc > 0;
            const bool open_leaf = has && open && nc == 0;
            // a node without children and without particles is the stub of a subtree that a locally
            // essential tree left out (bh_multigpu.cu): the sender's test guarantees that no target here
            // needs it opened, so it is accepted whatever this test says
            const bool pruned = FOREST && nc == 0 && nd.count == 0;
            const bool accept = has && (!open || pruned) && nd.cm.w != 0.f;
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
                if (VAR & 4) {
                    // transposed push: lane f of a round of 32 writes child number `base + f` of the round —
                    // consecutive words, no bank conflicts (a lane writing its own <= 8 children hits every
                    // bank incl / 32 times when its neighbours push 8 each) — and finds the parent that owns
                    // it by a shuffle binary search over the inclusive counts
                    for (int off = 0; off < total; off += 32) {
                        const int f = off + lane;
                        int owner = 0;
#pragma unroll
                        for (int step = 16; step >= 1; step >>= 1) {
                            const int v = __shfl_sync(FULL, incl, (owner + step - 1) & 31);
                            if (v <= f) owner += step;
                        }
                        owner = min(owner, 31);
                        const uint32_t ofc = __shfl_sync(FULL, nd.first_child, owner);
                        const int oex = __shfl_sync(FULL, incl - c_child, owner);
                        if (f < total) stack[sp + f] = ofc + (uint32_t)(f - oex);
                    }
                } else {
                    const int base = sp + incl - c_child;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < c_child) stack[base + j] = nd.first_child + j;
                }
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
        const bool acc = FOREST && a.accumulate;  // second phase of a two-phase walk: every row has one writer
        if (slice == 0 && tl < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ia] : (uint32_t)ia;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = acc ? o[0] + axa : axa;
            o[1] = acc ? o[1] + aya : aya;
            if (a.dim == 3) o[2] = acc ? o[2] + aza : aza;
        }
        if (slice == 0 && tl + half < gcnt) {
            const uint32_t row = a.tgt_perm ? a.tgt_perm[ib] : (uint32_t)ib;
            float *o = a.out + (size_t)row * a.dim;
            o[0] = acc ? o[0] + axb : axb;
            o[1] = acc ? o[1] + ayb : ayb;
            if (a.dim == 3) o[2] = acc ? o[2] + azb : azb;
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

// d_tgt == nullptr: the targets are the tree's own particles (the `&[P]` storage).
// tgt_stride: floats per target row (0 = bare positions, i.e. `dim`).

// Double precision (tree built by build64): d_tgt64 / d_out64 replace d_tgt / d_out; the f32 copy of
// separate targets that keys them is made here.
int traverse(pcuda_ctx *ctx, const pcuda_tree *t, const float *d_tgt, size_t na, float theta,
                    float eps, float *d_out, int tgt_stride, const double *d_tgt64, double *d_out64,
             double eps64) {
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
        launch_narrow(ctx, d_tgt64, na * ts, ctx->d_misc.as<float>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
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
        if (f64 && dim == 3) launch_gather64<3>(ctx, d_tgt64, ts, false, na, p, ctx->d_tgt_sorted.as<double4>());
        else if (f64) launch_gather64<2>(ctx, d_tgt64, ts, false, na, p, ctx->d_tgt_sorted.as<double4>());
        else if (dim == 3) launch_gather<3>(ctx, d_tgt, ts, false, na, p, ctx->d_tgt_sorted.as<float4>());
        else launch_gather<2>(ctx, d_tgt, ts, false, na, p, ctx->d_tgt_sorted.as<float4>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
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
// K5a from the TREE, when the targets are the tree's own particles and the one-pass build left the parent
// links: a segment (maximal cell with at most T targets) is then simply a node with at most T particles
// whose parent has more — every cell with more than T >= leaf_size particles is an internal node, so its
// children are nodes — and a run of equal keys longer than T is a leaf of the last level.  Same group
// starts as hard_flags + group_flags (0.41 ms at N = 10M: two neighbourhood searches per target), from
// one pass over the nodes (3.5M records).
__global__ void __launch_bounds__(256) group_flags_from_tree(const NodeRec *__restrict__ nodes,
                                                             const uint32_t *__restrict__ parent, uint32_t n_nodes,
                                                             uint32_t T, uint32_t gsize, uint32_t *__restrict__ flag) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const uint4 rec = reinterpret_cast<const uint4 *>(nodes + x)[1];  // first_child, nchild|level, begin, count
    const uint32_t begin = rec.z, count = rec.w;
    if (x != 0 && nodes[parent[x]].count <= T) return;  // inside a segment that starts higher up
    if (count <= T) {
        for (uint32_t i = begin; i < begin + count; i += gsize) flag[i] = 1u;
    } else if ((rec.y & 0xffu) == 0) {  // more than T targets with one key: cut at multiples of the group size
        flag[begin] = 1u;
        for (uint32_t i = (begin / gsize + 1) * gsize; i < begin + count; i += gsize) flag[i] = 1u;
    }
}

// K5a: the target groups of `n` targets from their keys, into the context's group buffers
// (ctx->d_stack, ctx->d_counters).  (Running this on a second stream beside the level build — both need
// only the sorted keys — was tried: build + walk 25.89 ms against 25.90 ms at N = 10M; the kernels of
// either side already fill the memory system, so nothing is hidden.)
// `own`: the tree whose particles the targets are (flags from its nodes), or nullptr (flags from the keys).
static int make_groups(pcuda_ctx *ctx, int dim, int bits, const uint64_t *tgt_keys, int n, int group_cap,
                       cudaStream_t st, const pcuda_tree *own = nullptr) {
    PCUDA_CUDA_TRY(ctx, ctx->d_counters.ensure(8 * sizeof(unsigned long long)));
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, 8 * sizeof(unsigned long long), st));
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
    if (own) {
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, (size_t)n * sizeof(uint32_t), st));
        group_flags_from_tree<<<(unsigned)((own->n_nodes + 255) / 256), 256, 0, st>>>(
            own->nodes.as<NodeRec>(), own->d_parent, (uint32_t)own->n_nodes, (uint32_t)g_seg_max, (uint32_t)group_cap,
            d_flag);
    } else {
        if (dim == 3) boundary_levels<3><<<nb256, 256, 0, st>>>(tgt_keys, n, d_L);
        else boundary_levels<2><<<nb256, 256, 0, st>>>(tgt_keys, n, d_L);
        const unsigned ngb = (unsigned)((n + GROUP_BLOCK - 1) / GROUP_BLOCK);
        hard_flags<<<ngb, GROUP_BLOCK, 0, st>>>(d_L, n, bits, g_seg_max, d_hard);
        group_flags<<<ngb, GROUP_BLOCK, 0, st>>>(d_hard, n, g_seg_max, group_cap, d_flag);
    }
    size_t tmp = 0;
    PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_flag, d_pos, n, st));
    PCUDA_CUDA_TRY(ctx, ctx->d_cub_tmp.ensure(tmp));
    PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub_tmp.p, tmp, d_flag, d_pos, n, st));
    scatter_groups<<<nb256, 256, 0, st>>>(d_flag, d_pos, n, d_gstart, d_ngroups);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += own ? 4 : 6;
    return PCUDA_OK;
}

int traverse_sorted(pcuda_ctx *ctx, const pcuda_tree *t, const float4 *tgt_sorted,
                           const uint64_t *tgt_keys, const uint32_t *tgt_perm, size_t na, float theta,
                           float eps, float *d_out, const Ext64 *x64, const ForestView *fv) {
    const int dim = t->dim;
    if (fv && (x64 || t->order == 2 || g_tpl != 2 || g_variant))
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT, "a forest is walked by traverse2_kernel only");
    cudaStream_t st = fv && fv->stream ? fv->stream : ctx->stream;
    const int group_cap = x64 ? 32 : 32 * g_tpl;  // the f64 walk holds one target per lane
    const int n = (int)na;
    // K5a: from the tree when the targets are its own particles (tuning hook bh_tree_groups = 0: from the keys)
    const bool own = g_tree_groups && t->d_parent && tgt_keys == t->d_keys() && na == t->n &&
                     (uint32_t)g_seg_max >= t->leaf_size;
    if (fv && fv->continue_groups) {
        // the groups the walk before did not get to (it was stopped): nothing to prepare
    } else if (fv && fv->reuse_groups) {  // same targets as the walk before: only the group dispenser starts over
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.as<unsigned long long>() + 3, 0, sizeof(uint32_t), st));
    } else {
        PCUDA_TRY(make_groups(ctx, dim, t->bits, tgt_keys, n, group_cap, st, own ? t : nullptr));
    }
    uint32_t *d_work = reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 3);
    uint32_t *d_ngroups = reinterpret_cast<uint32_t *>(ctx->d_counters.as<unsigned long long>() + 4);
    const size_t n4g = ((size_t)n + 3) & ~size_t(3);
    uint32_t *d_gstart = reinterpret_cast<uint32_t *>(ctx->d_stack.as<uint8_t>() + n4g) + 2 * (size_t)n;

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
    a.n_roots_dev = nullptr;
    a.accumulate = 0;
    a.reserve_from = a.reserve_cap = 0;
    a.chain = nullptr;
    a.stop = nullptr;
    a.share_scale = 1.f + 1.7320508f * theta;
    if (fv) {
        a.chain = fv->d_level_begin;
        a.stop = fv->d_stop;
        a.nodes = fv->nodes;
        a.src = fv->src;
        a.n_roots = fv->n_roots;
        a.roots = fv->d_roots;
        a.n_roots_dev = fv->d_n_roots;
        a.accumulate = fv->accumulate ? 1 : 0;
        if (fv->reserve_sms && fv->reserve_sms < (unsigned)ctx->sm_count) {
            a.reserve_from = (uint32_t)ctx->sm_count - fv->reserve_sms;
            a.reserve_cap = 4 * fv->reserve_sms;
        }
    }
    const size_t max_groups = ((size_t)n + 7) / 8;  // enough warps for small inputs, persistent beyond
    unsigned blocks = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 4,
                                                 (max_groups + TRAV_WARPS - 1) / TRAV_WARPS);
    if (blocks < (unsigned)ctx->sm_count * 4) a.reserve_cap = 0;  // a grid that does not fill the GPU leaves room anyway
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
        else if (g_variant == 3) traverse2_kernel<false, 3><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
        else traverse2_kernel<false, 4><<<blocks, TRAV_WARPS * 32, 0, st>>>(a);
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

int read_counters(pcuda_ctx *ctx) {
    if (!ctx->d_counters.p) return PCUDA_OK;
    unsigned long long h[6];  // [3], [4] hold the work dispenser and the group count
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_counters.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 3; ++i) ctx->last_counters[i] = h[i];
    ctx->last_counters[3] = h[5];
    ctx->last_counters[4] = h[4] & 0xffffffffull;  // number of target groups
    return PCUDA_OK;
}


}  // namespace bh
}  // namespace pcuda
