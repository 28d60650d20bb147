// bh_radix_build.cu — K4 in one pass: the linear orthtree over the sorted Morton keys, built from the
// boundary levels between neighbouring keys instead of level by level (Karras-style: every node is
// found from the key array alone, and the centres of mass climb bottom-up behind atomic arrival
// counters).  Seven launches whatever the depth of the tree; the level-wise build it replaces
// (expand_level / moments_kernel in barneshut.cu, kept for leaf sizes above RB_MAX_LEAF) needed
// 2 x (BITS + 1) dependent launches, each a round of dependent binary searches.
//
// The arrays are the ones of the tree specification (whose CPU statement the
// parity tests compare them with) bit for bit: nodes breadth-first, children of a node contiguous and in
// key order, moments in double precision added in key order (leaves) / child order (internal).
// The reference's own tree is the recursive bucket partition of particular/src/tree/mod.rs:91-138
// with the node data of gravity/impls/mod.rs:120-134.
//
//   L[i]   (i = 1 .. n-1) level at which keys i-1 and i fall into different cells (1 .. BITS;
//          BITS + 1 for equal keys);  L[0] = L[n] = 0.  Particle i starts a level-l cell iff L[i] <= l.
//   D[i]   level of the leaf that holds particle i = the smallest l whose cell around i has at most
//          `nleaf` particles = min over j <= i < k, k - j <= nleaf of max(L[j], L[k]), capped at BITS:
//          a window of nleaf keys to either side decides it.
//   nodes  = { (l, i) : L[i] <= l <= D[i] }: particle i starts one node on each of those levels (a
//          "chain"; the deepest one, (D[i], i), is a leaf).  Breadth-first index of (l, i) =
//          level_begin[l] + #{ i' < i : L[i'] <= l <= D[i'] }: per-tile counts per level (rb_count), an
//          exclusive scan over the tiles of every level (rb_scan), ranks inside the tile (rb_assign).
//   climb  one thread per leaf sums its particles and reports to the parent; the child that completes a
//          parent's arrivals sums the children's moments in child order and carries on upwards
//          (rb_moments).  A first child knows its parent (the node above it in its chain); the other
//          children find it through their left siblings (consecutive indices); the LAST child (the
//          next boundary after its cell is a boundary of the parent's level too) tells the parent how
//          many children it has, so no child count is needed in advance.
#include "bh.cuh"

namespace pcuda {
namespace bh {

constexpr int RB_TILE = 2048;            // particles per tile
constexpr int RB_BLOCK = 256;
constexpr int RB_ITERS = RB_TILE / RB_BLOCK;
constexpr int RB_CHUNKS = RB_TILE / 32;  // warp-sized runs of consecutive particles in a tile
constexpr int RB_LV = 33;                // level slots: 0 .. 32 (BITS <= 31)
constexpr uint32_t RB_NONE = 0xffffffffu;  // parent not known locally (not a first child)
constexpr uint32_t RB_ROOT = 0xfffffffeu;  // the root has no parent

template <int DIM>
__global__ void __launch_bounds__(256) rb_boundaries(const uint64_t *__restrict__ keys, uint32_t n,
                                                     uint8_t *__restrict__ L) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == 0 || i == n) {
        L[i] = 0;
        return;
    }
    const uint64_t x = keys[i] ^ keys[i - 1];
    L[i] = x == 0 ? (uint8_t)(Dims<DIM>::BITS + 1)
                  : (uint8_t)(Dims<DIM>::BITS - (63 - __clzll((long long)x)) / DIM);
}

// Leaf level of every particle + number of nodes every tile starts on every level.
// tile_cnt is level-major: tile_cnt[l * tiles_pad + tile].
template <int DIM>
__global__ void __launch_bounds__(RB_BLOCK) rb_count(const uint8_t *__restrict__ L, uint32_t n,
                                                     int nleaf, uint8_t *__restrict__ Dlv,
                                                     uint32_t *__restrict__ tile_cnt,
                                                     uint32_t tiles_pad) {
    constexpr int BITS = Dims<DIM>::BITS;
    constexpr int HALO = RB_MAX_LEAF;
    __shared__ uint8_t sL[RB_TILE + 2 * HALO];
    __shared__ uint32_t s_cnt[RB_LV];
    const long base = (long)blockIdx.x * RB_TILE;
    for (int q = threadIdx.x; q < RB_TILE + 2 * HALO; q += RB_BLOCK) {
        const long g = base - HALO + q;
        sL[q] = (g < 0 || g > (long)n) ? (uint8_t)255 : L[g];
    }
    if (threadIdx.x < RB_LV) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int it = 0; it < RB_ITERS; ++it) {
        const int off = it * RB_BLOCK + threadIdx.x;
        const long i = base + off;
        int li = 255, di = -1;  // an empty chain
        if (i < (long)n) {
            const uint8_t *c = sL + HALO + off;  // c[0] == L[i]
            int best = 255, r = 255;
            for (int b = 1; b <= nleaf; ++b) {
                r = min(r, (int)c[b]);
                best = min(best, max((int)c[b - nleaf], r));
            }
            di = min(best, BITS);
            li = c[0];
            Dlv[i] = (uint8_t)di;
        }
        const bool chain = li <= di;
        int lo = chain ? li : 255, hi = chain ? di : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL, lo, o));
            hi = max(hi, __shfl_xor_sync(FULL, hi, o));
        }
        for (int l = lo; l <= hi; ++l) {
            const unsigned m = __ballot_sync(FULL, chain && li <= l && l <= di);
            if (lane == 0 && m) atomicAdd(&s_cnt[l], (uint32_t)__popc(m));
        }
    }
    __syncthreads();
    if (threadIdx.x < RB_LV) tile_cnt[(size_t)threadIdx.x * tiles_pad + blockIdx.x] = s_cnt[threadIdx.x];
}

// Exclusive scan over the tiles of every level (in place) + the level table.  One block per level; the
// block that finishes last (a ticket) turns the level totals into the level table.  (One block for all
// levels took 57 us at N = 10M: 4883 tiles x 33 levels behind one another.)
__global__ void __launch_bounds__(1024) rb_scan(uint32_t *__restrict__ tile_cnt, uint32_t n_tiles,
                                                uint32_t tiles_pad, BuildState *st, uint32_t capacity,
                                                uint32_t *__restrict__ totals /* RB_LV + 1: totals, ticket */) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running;
    __shared__ bool s_last;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l = blockIdx.x;
    uint32_t *row = tile_cnt + (size_t)l * tiles_pad;
    if (threadIdx.x == 0) s_running = 0;
    __syncthreads();
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += 1024) {
        const uint32_t t = t0 + threadIdx.x;
        const uint32_t v = t < n_tiles ? row[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;  // inclusive over the warps
        }
        __syncthreads();
        const uint32_t base = s_running + (warp ? s_warp[warp - 1] : 0u);
        if (t < n_tiles) row[t] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_running += s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[l] = s_running;
        __threadfence();
        s_last = atomicAdd(&totals[RB_LV], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        unsigned long long b = 0;
        for (int k = 0; k < 36; ++k) {
            st->level_begin[k] = b > 0xffffffffull ? 0xffffffffu : (uint32_t)b;
            if (k < RB_LV) b += __ldcg(&totals[k]);
        }
        for (int k = 0; k < 34; ++k) st->ticket[k] = 0;
        st->overflow = b > capacity ? 1u : 0u;
        st->capacity = capacity;
        totals[RB_LV] = 0;  // ready for the next build
    }
}

// The node records (structure only), the parent links of first children and the arrival counters.
template <int DIM>
__global__ void __launch_bounds__(RB_BLOCK) rb_assign(const uint8_t *__restrict__ L,
                                                      const uint8_t *__restrict__ Dlv, uint32_t n,
                                                      const uint32_t *__restrict__ tile_base,
                                                      uint32_t tiles_pad,
                                                      const BuildState *__restrict__ st,
                                                      NodeRec *__restrict__ nodes,
                                                      uint32_t *__restrict__ parent,
                                                      uint32_t *__restrict__ arrive) {
    __shared__ uint32_t s_cnt[RB_CHUNKS][RB_LV];
    if (st->overflow) return;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long base = (long)blockIdx.x * RB_TILE;
    int li[RB_ITERS], di[RB_ITERS];
    for (int q = threadIdx.x; q < RB_CHUNKS * RB_LV; q += RB_BLOCK) (&s_cnt[0][0])[q] = 0;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RB_ITERS; ++it) {
        const long i = base + it * RB_BLOCK + threadIdx.x;
        li[it] = 255;
        di[it] = -1;
        if (i < (long)n) {
            li[it] = L[i];
            di[it] = Dlv[i];
        }
        const bool chain = li[it] <= di[it];
        int lo = chain ? li[it] : 255, hi = chain ? di[it] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL, lo, o));
            hi = max(hi, __shfl_xor_sync(FULL, hi, o));
        }
        const int chunk = it * (RB_BLOCK / 32) + warp;
        for (int l = lo; l <= hi; ++l) {
            const unsigned m = __ballot_sync(FULL, chain && li[it] <= l && l <= di[it]);
            if (lane == 0) s_cnt[chunk][l] = (uint32_t)__popc(m);
        }
    }
    __syncthreads();
    if (threadIdx.x < RB_LV) {  // index of the first node every chunk starts on level l
        const int l = threadIdx.x;
        uint32_t running = st->level_begin[l] + tile_base[(size_t)l * tiles_pad + blockIdx.x];
        for (int c = 0; c < RB_CHUNKS; ++c) {
            const uint32_t v = s_cnt[c][l];
            s_cnt[c][l] = running;
            running += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RB_ITERS; ++it) {
        const uint32_t i = (uint32_t)(base + it * RB_BLOCK + threadIdx.x);
        const bool chain = li[it] <= di[it];
        int lo = chain ? li[it] : 255, hi = chain ? di[it] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(FULL, lo, o));
            hi = max(hi, __shfl_xor_sync(FULL, hi, o));
        }
        const int chunk = it * (RB_BLOCK / 32) + warp;
        uint32_t prev = RB_NONE;
        for (int l = lo; l <= hi; ++l) {
            const bool active = chain && li[it] <= l && l <= di[it];
            const unsigned m = __ballot_sync(FULL, active);
            if (!active) continue;
            const uint32_t id = s_cnt[chunk][l] + (uint32_t)__popc(m & ((1u << lane) - 1u));
            arrive[id] = 0;
            if (l > li[it]) {  // the node above in the chain is internal and this is its first child
                parent[id] = prev;
                uint4 rec = make_uint4(id, (uint32_t)(l - 1) << 8, i, 0u);
                reinterpret_cast<uint4 *>(nodes + prev)[1] = rec;
            } else {
                parent[id] = l == 0 ? RB_ROOT : RB_NONE;
            }
            if (l == di[it]) {  // the leaf of the chain: its cell ends at the next boundary of its level
                uint32_t k = i + 1;
                while ((k & 3u) && L[k] > l) ++k;
                if (!(k & 3u) && L[k] > l) {
                    const uint32_t lim = (uint32_t)l * 0x01010101u;
                    for (;;) {
                        const uint32_t w = *reinterpret_cast<const uint32_t *>(L + k);
                        const uint32_t le = __vcmpleu4(w, lim);  // 0xff in every byte with L <= l
                        if (le) {
                            k += (uint32_t)(__ffs((int)le) - 1) >> 3;
                            break;
                        }
                        k += 4;
                    }
                }
                uint4 rec = make_uint4(0u, (uint32_t)l << 8, i, k - i);
                reinterpret_cast<uint4 *>(nodes + id)[1] = rec;
            }
            prev = id;
        }
    }
}

// Links of every node to its parent, so that the climb does not have to search: a first child knows
// its parent from rb_assign; the other children find the first one among their left siblings
// (consecutive indices, at most 2^DIM - 1 steps).  A child is the LAST one when its right neighbour
// is a first child (the first node of a level is one) or does not exist; it then knows how many
// children the parent has, writes that into the parent's record and pre-loads the parent's arrival
// counter with it (bits 8 and up), so that the climb only counts arrivals (bits 0..7) until both agree.
__global__ void __launch_bounds__(256) rb_links(NodeRec *__restrict__ nodes,
                                                const uint32_t *__restrict__ parent,
                                                uint32_t *__restrict__ plink,
                                                uint32_t *__restrict__ arrive,
                                                const BuildState *__restrict__ st) {
    if (st->overflow) return;
    const uint32_t n_nodes = st->level_begin[35];
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    uint32_t p = parent[x], s = 0;
    while (p == RB_NONE) {
        ++s;
        p = parent[x - s];
    }
    plink[x] = p;
    if (p == RB_ROOT) return;
    if (x + 1 == n_nodes || parent[x + 1] != RB_NONE) {
        const uint32_t nc = s + 1, level = nodes[x].nchild_level >> 8;
        nodes[p].nchild_level = nc | (level - 1) << 8;
        arrive[p] = nc << 8;
    }
}

// Moments and centres of mass, bottom-up (see the header comment).  Arithmetic and order of
// node_moments() in bh_build.cu == the CPU statement of the specification.
template <int DIM>
__device__ __forceinline__ void rb_store(NodeRec *__restrict__ nodes, double *__restrict__ mom,
                                         const float4 *__restrict__ sorted, uint32_t x, const double m[4],
                                         uint32_t beg) {
    reinterpret_cast<double4 *>(mom)[x] = make_double4(m[0], m[1], m[2], m[3]);
    float4 cm;
    if (m[3] == 0.0) {
        const float4 q = sorted[beg];
        cm = make_float4(q.x, q.y, DIM == 3 ? q.z : 0.f, 0.f);
    } else {
        cm.x = (float)__ddiv_rn(m[0], m[3]);
        cm.y = (float)__ddiv_rn(m[1], m[3]);
        cm.z = DIM == 3 ? (float)__ddiv_rn(m[2], m[3]) : 0.f;
        cm.w = (float)m[3];
    }
    nodes[x].cm = cm;
}

// Pass 1, no synchronisation at all: every leaf sums its particles (key order); every internal node notes
// whether all its children are leaves.
template <int DIM>
__global__ void __launch_bounds__(128) rb_leaf_moments(NodeRec *__restrict__ nodes, double *__restrict__ mom,
                                                       const float4 *__restrict__ sorted,
                                                       uint8_t *__restrict__ all_leaf,
                                                       const BuildState *__restrict__ st) {
    if (st->overflow) return;
    const uint32_t n_nodes = st->level_begin[35];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_nodes; j += gridDim.x * blockDim.x) {
        const uint4 rec = reinterpret_cast<const uint4 *>(nodes + j)[1];
        if (rec.x != 0) {
            bool all = true;
            for (uint32_t c = 0; c < (rec.y & 0xffu); ++c) all &= nodes[rec.x + c].first_child == 0;
            all_leaf[j] = all ? 1 : 0;
            continue;
        }
        const uint32_t beg = rec.z, end = rec.z + rec.w;
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        for (uint32_t i = beg; i < end; i += 4) {  // four loads in flight, additions in key order
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < end) q[u] = sorted[i + u];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i + u < end) {
                    const double mi = (double)q[u].w;
                    m[0] = __dadd_rn(m[0], __dmul_rn(mi, (double)q[u].x));
                    m[1] = __dadd_rn(m[1], __dmul_rn(mi, (double)q[u].y));
                    if (DIM == 3) m[2] = __dadd_rn(m[2], __dmul_rn(mi, (double)q[u].z));
                    m[3] = __dadd_rn(m[3], mi);
                }
            }
        }
        rb_store<DIM>(nodes, mom, sorted, j, m, beg);
    }
}

// Pass 2, the climb.  An internal node whose children are all leaves (most of the internal nodes) sums
// them directly — their moments were written by pass 1, a kernel ago — so those leaves never touch an
// arrival counter; the other leaves only count themselves in.  A node that was computed HERE is released
// before its arrival is counted, and the child that completes a count acquires and reads its siblings from
// L2 (ld.cg).  (One kernel for everything, every leaf arriving with a fence: 397 us at N = 10M; with
// __threadfence() pairs and the parent searched in the climb: 491 us.)
template <int DIM>
__global__ void __launch_bounds__(128) rb_climb(NodeRec *__restrict__ nodes, double *__restrict__ mom,
                                                const float4 *__restrict__ sorted,
                                                const uint32_t *__restrict__ plink, uint32_t *__restrict__ arrive,
                                                const uint8_t *__restrict__ all_leaf,
                                                const BuildState *__restrict__ st) {
    if (st->overflow) return;
    const uint32_t n_nodes = st->level_begin[35];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_nodes; j += gridDim.x * blockDim.x) {
        const uint4 rec = reinterpret_cast<const uint4 *>(nodes + j)[1];
        uint32_t x = j, p = plink[j];
        bool wrote = false;  // x's data was written by this thread in this kernel: release before arriving
        if (rec.x == 0) {
            if (p == RB_ROOT || all_leaf[p]) continue;  // the parent sums this leaf itself
        } else if (all_leaf[j]) {
            double m[4] = {0.0, 0.0, 0.0, 0.0};
            uint32_t total = 0;
            for (uint32_t c = 0; c < (rec.y & 0xffu); ++c) {
                const double4 q = reinterpret_cast<const double4 *>(mom)[rec.x + c];
                m[0] = __dadd_rn(m[0], q.x);
                m[1] = __dadd_rn(m[1], q.y);
                if (DIM == 3) m[2] = __dadd_rn(m[2], q.z);
                m[3] = __dadd_rn(m[3], q.w);
                total += nodes[rec.x + c].count;
            }
            rb_store<DIM>(nodes, mom, sorted, j, m, rec.z);
            nodes[j].count = total;
            wrote = true;
        } else {
            continue;  // completed by the child that arrives last
        }
        while (p != RB_ROOT) {
            uint32_t now;
            if (wrote) asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(now) : "l"(arrive + p) : "memory");
            else now = atomicAdd(arrive + p, 1u);
            now += 1u;
            if ((now & 0xffu) != (now >> 8)) break;  // siblings still on their way
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            const uint32_t nc = now >> 8;
            const uint4 prec = reinterpret_cast<const uint4 *>(nodes + p)[1];  // written by earlier kernels
            const uint32_t fc = prec.x;
            double m[4] = {0.0, 0.0, 0.0, 0.0};
            uint32_t total = 0;
            for (uint32_t c = 0; c < nc; ++c) {
                const double2 a = __ldcg(reinterpret_cast<const double2 *>(mom) + 2 * (size_t)(fc + c));
                const double2 b = __ldcg(reinterpret_cast<const double2 *>(mom) + 2 * (size_t)(fc + c) + 1);
                m[0] = __dadd_rn(m[0], a.x);
                m[1] = __dadd_rn(m[1], a.y);
                if (DIM == 3) m[2] = __dadd_rn(m[2], b.x);
                m[3] = __dadd_rn(m[3], b.y);
                total += __ldcg(&nodes[fc + c].count);
            }
            x = p;
            rb_store<DIM>(nodes, mom, sorted, x, m, prec.z);
            nodes[x].count = total;
            wrote = true;
            p = plink[x];
        }
    }
}

template <int DIM>
int radix_build_enqueue(pcuda_ctx *ctx, pcuda_tree *t, size_t n, size_t cap_nodes, BuildState *d_state) {
    cudaStream_t st = ctx->stream;
    const uint32_t n_tiles = (uint32_t)((n + RB_TILE - 1) / RB_TILE);
    const uint32_t tiles_pad = (n_tiles + 31u) & ~31u;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t off_L = 0, off_D = up(n + 1 + 4), off_cnt = off_D + up(n),
                 off_parent = off_cnt + up((size_t)RB_LV * tiles_pad * 4),
                 off_plink = off_parent + up(cap_nodes * 4), off_arrive = off_plink + up(cap_nodes * 4),
                 off_leaf = off_arrive + up(cap_nodes * 4), off_totals = off_leaf + up(cap_nodes),
                 total = off_totals + up((RB_LV + 1) * 4);
    const bool fresh = total > t->rb.cap;  // a new allocation: the scan's ticket must start at zero
    PCUDA_CUDA_TRY(ctx, t->rb.ensure(total));
    uint8_t *base = t->rb.as<uint8_t>();
    uint8_t *L = base + off_L, *Dlv = base + off_D;
    uint32_t *tile_cnt = reinterpret_cast<uint32_t *>(base + off_cnt);
    uint32_t *parent = reinterpret_cast<uint32_t *>(base + off_parent);
    uint32_t *plink = reinterpret_cast<uint32_t *>(base + off_plink);
    uint32_t *arrive = reinterpret_cast<uint32_t *>(base + off_arrive);
    uint8_t *all_leaf = base + off_leaf;
    uint32_t *totals = reinterpret_cast<uint32_t *>(base + off_totals);
    if (fresh || t->rb_totals != totals) {
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(totals, 0, (RB_LV + 1) * 4, st));
        t->rb_totals = totals;
    }
    const int nleaf = (int)t->leaf_size;
    NodeRec *nodes = t->nodes.as<NodeRec>();
    rb_boundaries<DIM><<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(t->d_keys(), (uint32_t)n, L);
    rb_count<DIM><<<n_tiles, RB_BLOCK, 0, st>>>(L, (uint32_t)n, nleaf, Dlv, tile_cnt, tiles_pad);
    rb_scan<<<RB_LV, 1024, 0, st>>>(tile_cnt, n_tiles, tiles_pad, d_state, (uint32_t)cap_nodes, totals);
    rb_assign<DIM><<<n_tiles, RB_BLOCK, 0, st>>>(L, Dlv, (uint32_t)n, tile_cnt, tiles_pad, d_state, nodes,
                                                 parent, arrive);
    rb_links<<<(unsigned)((cap_nodes + 255) / 256), 256, 0, st>>>(nodes, parent, plink, arrive, d_state);
    const unsigned mgrid = (unsigned)std::min<size_t>((size_t)ctx->sm_count * 16, (cap_nodes + 127) / 128);
    rb_leaf_moments<DIM><<<mgrid, 128, 0, st>>>(nodes, t->moments.as<double>(), t->sorted.as<float4>(), all_leaf,
                                                d_state);
    rb_climb<DIM><<<mgrid, 128, 0, st>>>(nodes, t->moments.as<double>(), t->sorted.as<float4>(), plink, arrive,
                                         all_leaf, d_state);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 7;
    t->d_parent = plink;  // (the root: RB_ROOT = 0xfffffffe)
    return PCUDA_OK;
}

template int radix_build_enqueue<2>(pcuda_ctx *, pcuda_tree *, size_t, size_t, BuildState *);
template int radix_build_enqueue<3>(pcuda_ctx *, pcuda_tree *, size_t, size_t, BuildState *);

}  // namespace bh
}  // namespace pcuda
