// bh_multigpu.cu — multi-GPU Barnes-Hut (one process per GPU): the replicated build, the key-range
// partitioned build joined by a top tree, the routing of the accelerations to the ranks that own the
// particles, and the single-GPU "virtual rank" test entry.  New functionality: the reference is
// single-device.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

#include "bh.cuh"

namespace pcuda {
namespace bh {

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
    // locally essential trees (sharded_let_dev)
    pcuda::DevBuf let_boxes, let_box_all, let_keys[2], let_idx[2], let_send_rec, let_recv_rec,
        let_cuts, let_cnt_mat, let_dom, let_dom_all, let_open, let_reach, let_parent,
        let_tile_cnt, let_totals, let_index, let_send_nodes, let_send_src, let_bmap_send, let_bmap_recv, let_gi;
    uint32_t *h_let = nullptr;  // pinned: count matrices and boundary maps
    // two-phase walk: the rank's own tree is walked on walk_stream while the others' trees are on their way
    pcuda::DevBuf let_root0;
    cudaStream_t walk_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace pcuda {

void forest_free(pcuda_ctx *ctx) {
    pcuda_forest *f = ctx->forest;
    if (!f) return;
    if (f->local) tree_free(ctx, f->local);
    DevBuf *bufs[] = {&f->gkeys, &f->gidx, &f->sample[0], &f->sample[1], &f->split, &f->counts,
                      &f->sel_tmp, &f->sel_count, &f->nodes, &f->sorted, &f->perm, &f->keys, &f->acc,
                      &f->packs, &f->stage, &f->roots, &f->route_cnt, &f->route_pos, &f->route_idx_send,
                      &f->route_acc_send, &f->route_idx_recv, &f->route_acc_recv, &f->let_boxes, &f->let_box_all,
                      &f->let_keys[0], &f->let_keys[1], &f->let_idx[0], &f->let_idx[1], &f->let_send_rec,
                      &f->let_recv_rec, &f->let_cuts,
                      &f->let_cnt_mat, &f->let_dom, &f->let_dom_all, &f->let_open, &f->let_reach, &f->let_parent,
                      &f->let_tile_cnt, &f->let_totals, &f->let_index, &f->let_send_nodes, &f->let_send_src,
                      &f->let_bmap_send, &f->let_bmap_recv, &f->let_gi, &f->let_root0};
    for (DevBuf *b : bufs) b->release();
    if (f->h_let) cudaFreeHost(f->h_let);
    if (f->h_route) cudaFreeHost(f->h_route);
    if (f->h_packs) cudaFreeHost(f->h_packs);
    if (f->h_stage) cudaFreeHost(f->h_stage);
    if (f->ev_stage) cudaEventDestroy(f->ev_stage);
    if (f->ev_fork) cudaEventDestroy(f->ev_fork);
    if (f->ev_join) cudaEventDestroy(f->ev_join);
    if (f->walk_stream) cudaStreamDestroy(f->walk_stream);
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
    launch_encode<3>(ctx, d_particles, 4, n, t->d_frame.as<Frame>(), f->gkeys.as<uint64_t>(), f->gidx.as<uint32_t>());
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
    ctx->launches += 4 + 9;
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
        launch_gather<3>(ctx, d_particles, 4, true, count, t->d_perm(), t->sorted.as<float4>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 2 + 1 + 9;
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
// gi_table (optional, [part][level][side]): index of every boundary node in the joined array when the
// parts are not stored with their own numbering (locally essential trees: pruned copies).
static int merge_top_tree(pcuda_ctx *ctx, int parts, const PartPack *packs, const BoundaryRec *stage,
                          const uint32_t *node_base, uint32_t top_base, std::vector<NodeRec> &top,
                          std::vector<uint32_t> &roots, const uint32_t *gi_table = nullptr) {
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
                in.gi = gi_table ? gi_table[((size_t)q * TOP_LEVELS + l) * 2 + side]
                                 : node_base[q] + (side ? le - 1 : lb);
                // a first / last node of a level that the locally essential tree does not hold lies below a
                // pruned node: a complete cell nobody will ask for (the chain of partial cells is always sent)
                if (gi_table && in.gi == 0xffffffffu) continue;
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
        r.nchild_level = (uint32_t)cells[c].l << 8 | (first.nchild_level & NODE_SHARE);
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
            r.nchild_level = (uint32_t)cur.level << 8 | (top[cur.me].nchild_level & NODE_SHARE);  // a share's rest is a share
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
        top[cur.me].nchild_level = (top[cur.me].nchild_level & NODE_SHARE) | (uint32_t)cur.level << 8 | (uint32_t)cur.kids.size();
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
int partitioned_dev(pcuda_ctx *ctx, const float *d_particles, size_t n, int parts, float theta,
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

// ------------------------------------------------------------------------------------------------
// Locally essential trees (the replicating exchanges of the two builds above are what
// limits them: every rank receives all N records, all nodes and all accelerations).  Here nothing is
// replicated:
//
//   A. the particles go to the rank that owns their KEY RANGE: common root cube from the ranks' boxes
//      (all-gather of 32 B), keys of the local block, local sort, splitters from an all-gathered sample of
//      the sorted keys, one all-to-all of the records (each travels once: 16 B x N / world per rank
//      instead of 16 B x N; no index travels, see E);
//   B. every rank sorts what it received (equal keys end up in global-index order, as on one GPU) and
//      builds the tree of its range in the common cube (one-pass build);
//   C. every rank tells the others WHERE its targets are — the top of its tree as a tree of boxes whose
//      frontier cells each hold whole target groups of the walk (LetDomain) — and sends each of them
//      only what their walk can touch: node x goes to rank q when all its ancestors are opened by the
//      opening rule against some frontier cell of q (the rule of the walk, made conservative by a
//      margin), the particles of a leaf when the leaf itself is.  Open flags for every (node, rank) pair, an AND along the ancestors, a
//      compaction per destination that keeps breadth-first order (children stay contiguous): five small
//      data-parallel kernels, no work queues.  A pruned node travels as a record without children or
//      particles, which the walk accepts whatever its own test says;
//   D. cells that straddle a range boundary are joined into the top tree exactly as in the partitioned
//      build (merge_top_tree), from the boundary nodes that every locally essential tree always carries;
//   E. the walk starts at the top root and writes every acceleration where its particle arrived in A;
//      the blocks return to their senders with the counts of A swapped (12 B x N / world per rank) and
//      each owner scatters them through its own sort permutation.
//
// Exchange per rank at N = 10M on 8 GPUs: 25 + ~25 + 20 MB instead of 140 + 140 + 105 MB.
constexpr int LET_SAMPLE = 4096;  // key samples per rank
constexpr int LET_SPLIT_BITS = 32;  // leading key bits that decide the destination of a particle (10.7 levels)
constexpr int LET_TILE = 1024;    // nodes per block of the compaction kernels

struct LetTotals {  // per destination: counts and send offsets of nodes / particles (device + host copy)
    uint32_t n_nodes[MAX_PARTS], n_src[MAX_PARTS], off_nodes[MAX_PARTS], off_src[MAX_PARTS];
};

// Where the targets of a rank are: the top of its tree as a tree of boxes, refined while a cell holds more
// than tau = max(seg_max, 8 n / LET_DOM_MAX) particles.  A FRONTIER cell (nchild == 0) is a leaf or a cell
// with at most tau particles — only cells with more than seg_max particles are split, so every boundary
// between two frontier cells is a "hard"
// boundary of the target grouping (hard_flags: the smallest cell holding both neighbours has more than
// seg_max targets) and every target group of the walk lies inside ONE frontier cell.  The box of a cell
// is its cube (from the key prefix), widened by a margin that covers the rounding of the quantisation.
// (A first version used the boxes of 64 windows of consecutive particles: a window that straddles a jump
// of the Z-order curve has a box spanning half the cloud, and one rank received the other's whole tree.)
constexpr int LET_DOM_MAX = 4096;
struct LetDomNode {
    float lo[3], hi[3];
    uint32_t first_child, nchild;  // nchild == 0: frontier cell
};
struct LetDomain {
    uint32_t n;  // 0: the rank has no targets
    uint32_t pad;
    float lo[3], hi[3];  // box of all frontier cells: one test rules out most nodes of a distant rank
    LetDomNode node[LET_DOM_MAX];
};

// parent[x] of every node (the root: 0xffffffff), from the child ranges.
__global__ void __launch_bounds__(256) let_parent_kernel(const NodeRec *__restrict__ nodes, uint32_t n_nodes,
                                                         uint32_t *__restrict__ parent) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    if (x == 0) parent[0] = 0xffffffffu;
    const NodeRec nd = nodes[x];
    for (uint32_t j = 0; j < (nd.nchild_level & 0xffu); ++j) parent[nd.first_child + j] = x;
}

// The domain tree = the nodes whose parent holds more than `tau` particles (the root included): cells are
// split while they are heavy, so the frontier is fine where the targets are dense.  flag -> (scan) ->
// idx keeps breadth-first order, hence the children of a split cell stay contiguous.
__global__ void __launch_bounds__(256) let_dom_flag_kernel(const NodeRec *__restrict__ nodes,
                                                           const uint32_t *__restrict__ parent, uint32_t n_nodes,
                                                           uint32_t tau, uint32_t *__restrict__ flag) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    flag[x] = (x == 0 || nodes[parent[x]].count > tau) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) let_domain_kernel(const NodeRec *__restrict__ nodes,
                                                         const uint64_t *__restrict__ keys,
                                                         const uint32_t *__restrict__ flag,
                                                         const uint32_t *__restrict__ idx, uint32_t n_nodes,
                                                         uint32_t tau, const Frame *__restrict__ frame,
                                                         LetDomain *__restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    if (x == n_nodes - 1) out->n = min(idx[x] + flag[x], (uint32_t)LET_DOM_MAX);
    if (!flag[x]) return;
    const uint32_t y = idx[x];
    if (y >= (uint32_t)LET_DOM_MAX) return;  // table full: the parent becomes a frontier cell (below)
    const NodeRec nd = nodes[x];
    const uint32_t nc = nd.nchild_level & 0xffu, level = nd.nchild_level >> 8;
    const uint64_t pre = level ? keys[nd.begin] >> (3 * (Dims<3>::BITS - level)) : 0ull;
    uint32_t q[3] = {0, 0, 0};  // cell coordinates: de-interleave the prefix (axis 0 in the lowest bit)
    for (uint32_t b = 0; b < level; ++b)
        for (int k = 0; k < 3; ++k) q[k] |= (uint32_t)((pre >> (3 * b + k)) & 1ull) << b;
    const float ext = frame->ext;
    const float w = ext * __int_as_float((127 - (int)level) << 23);
    const float margin = ext * 2e-6f;  // > the rounding of (x - origin) * inv and of the corner below
    LetDomNode o;
    for (int k = 0; k < 3; ++k) {
        const float c = frame->origin[k] + (float)q[k] * w;
        o.lo[k] = c - margin;
        o.hi[k] = c + w + margin;
    }
    const bool split = nc > 0 && nd.count > tau && idx[nd.first_child] + nc <= (uint32_t)LET_DOM_MAX;
    o.first_child = split ? idx[nd.first_child] : 0u;
    o.nchild = split ? nc : 0u;
    out->node[y] = o;
}

// Box of the whole domain = union of the cubes of the domain tree's leaves (one block).
__global__ void __launch_bounds__(256) let_domain_box_kernel(LetDomain *d) {
    __shared__ float s[8][6];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t y = threadIdx.x; y < d->n; y += 256) {
        const LetDomNode &nd = d->node[y];
        if (nd.nchild) continue;
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], nd.lo[k]);
            hi[k] = fmaxf(hi[k], nd.hi[k]);
        }
    }
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; ++k) {
            s[threadIdx.x >> 5][k] = lo[k];
            s[threadIdx.x >> 5][3 + k] = hi[k];
        }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = s[0][threadIdx.x], b = s[0][3 + threadIdx.x];
        for (int w = 1; w < 8; ++w) {
            a = fminf(a, s[w][threadIdx.x]);
            b = fmaxf(b, s[w][3 + threadIdx.x]);
        }
        d->lo[threadIdx.x] = a;
        d->hi[threadIdx.x] = b;
    }
}

// The opening rule of the walk (traverse2_kernel: theta^2 dmin^2 < w^2, dmin = distance from the
// centre of mass to the box of the targets) against a box that CONTAINS the box of every target group
// it stands for.  `margin` (a few ulp of the root extent) is taken off every axis distance, so that the
// different rounding of the walk's centre / half-width form can never make the walk open a node that
// this test kept closed.
__device__ __forceinline__ bool let_opens(const float4 cm, float w2, float theta2, const float *lo,
                                          const float *hi, float margin) {
    const float dx = fmaxf(fmaxf(lo[0] - cm.x, cm.x - hi[0]) - margin, 0.f);
    const float dy = fmaxf(fmaxf(lo[1] - cm.y, cm.y - hi[1]) - margin, 0.f);
    const float dz = fmaxf(fmaxf(lo[2] - cm.z, cm.z - hi[2]) - margin, 0.f);
    const float d2 = dx * dx + dy * dy + dz * dz;
    return theta2 * d2 < w2;
}

// open[x] bit q: rank q's walk may open node x (an internal node: it needs the children; a leaf: it
// needs the particles).  Boundary nodes (first / last of a level: the only cells that other ranks may
// hold a share of) are always open: the walk tests the MERGED cell, whose centre of mass this rank
// does not know.
__global__ void __launch_bounds__(256) let_open_kernel(const NodeRec *__restrict__ nodes, uint32_t n_nodes,
                                                       const BuildState *__restrict__ st,
                                                       const LetDomain *__restrict__ doms, int world, int rank,
                                                       float theta2, const Frame *__restrict__ frame,
                                                       uint16_t *__restrict__ open) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const NodeRec nd = nodes[x];
    const uint32_t nc = nd.nchild_level & 0xffu, level = nd.nchild_level >> 8;
    const float ext = frame->ext;
    const float w = ext * __int_as_float((127 - (int)level) << 23);
    const float w2 = w * w, margin = ext * 1e-6f;
    const bool boundary = nc > 0 && (x == st->level_begin[level] || x + 1 == st->level_begin[level + 1]);
    uint32_t mask = 0;
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        const LetDomain &d = doms[q];
        bool op = boundary;
        // walk q's domain tree: a box that passes the rule closes its whole branch (first of all the box of
        // the whole domain)
        if (!op && d.n && let_opens(nd.cm, w2, theta2, d.lo, d.hi, margin)) {
            uint16_t stack[192];  // <= 7 siblings per level of the domain tree + 8
            int sp = 1;
            stack[0] = 0;
            while (sp && !op) {
                const LetDomNode &y = d.node[stack[--sp]];
                if (!let_opens(nd.cm, w2, theta2, y.lo, y.hi, margin)) continue;
                if (y.nchild == 0) op = true;
                else
                    for (uint32_t j = 0; j < y.nchild; ++j) stack[sp++] = (uint16_t)(y.first_child + j);
            }
        }
        mask |= (uint32_t)op << q;
    }
    open[x] = (uint16_t)mask;
}

// reach[x] bit q: every ancestor of x is open for q, i.e. x belongs to the tree sent to q.
__global__ void __launch_bounds__(256) let_reach_kernel(const uint16_t *__restrict__ open,
                                                        const uint32_t *__restrict__ parent, uint32_t n_nodes,
                                                        uint32_t all, uint16_t *__restrict__ reach) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    uint32_t m = all, a = parent[x];
    while (a != 0xffffffffu && m) {
        m &= open[a];
        a = parent[a];
    }
    reach[x] = (uint16_t)m;
}

// Per tile and destination: nodes sent, particles sent.  tile_cnt[(2 q + k) * tiles_pad + tile].
__global__ void __launch_bounds__(256) let_count_kernel(const NodeRec *__restrict__ nodes, uint32_t n_nodes,
                                                        const uint16_t *__restrict__ open,
                                                        const uint16_t *__restrict__ reach, int world,
                                                        uint32_t *__restrict__ tile_cnt, uint32_t tiles_pad) {
    __shared__ uint32_t s_cnt[2 * MAX_PARTS];
    if (threadIdx.x < 2 * MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    uint32_t cn[MAX_PARTS], cp[MAX_PARTS];
#pragma unroll
    for (int q = 0; q < MAX_PARTS; ++q) cn[q] = cp[q] = 0;
    for (int u = 0; u < LET_TILE / 256; ++u) {
        const uint32_t x = blockIdx.x * LET_TILE + threadIdx.x * (LET_TILE / 256) + u;
        if (x >= n_nodes) break;
        const uint32_t r = reach[x];
        if (!r) continue;
        const NodeRec nd = nodes[x];
        const uint32_t send_p = (nd.nchild_level & 0xffu) == 0 ? (r & open[x]) : 0u;
#pragma unroll
        for (int q = 0; q < MAX_PARTS; ++q) {
            cn[q] += (r >> q) & 1u;
            cp[q] += ((send_p >> q) & 1u) ? nd.count : 0u;
        }
    }
#pragma unroll
    for (int q = 0; q < MAX_PARTS; ++q) {
        if (q >= world) break;
        uint32_t a = cn[q], b = cp[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (a) atomicAdd(&s_cnt[2 * q], a);
            if (b) atomicAdd(&s_cnt[2 * q + 1], b);
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < 2 * world) tile_cnt[(size_t)threadIdx.x * tiles_pad + blockIdx.x] = s_cnt[threadIdx.x];
}

// Exclusive scan over the tiles of every (destination, kind) row, totals and send offsets.
__global__ void __launch_bounds__(1024) let_scan_kernel(uint32_t *__restrict__ tile_cnt, uint32_t n_tiles,
                                                        uint32_t tiles_pad, int world, LetTotals *__restrict__ tot) {
    __shared__ uint32_t s_total[2 * MAX_PARTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 2 * world) {
        uint32_t *row = tile_cnt + (size_t)warp * tiles_pad;
        uint32_t running = 0;
        for (uint32_t t0 = 0; t0 < n_tiles; t0 += 32) {
            const uint32_t t = t0 + lane;
            const uint32_t v = t < n_tiles ? row[t] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (t < n_tiles) row[t] = running + incl - v;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_total[warp] = running;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t on = 0, os = 0;
        for (int q = 0; q < MAX_PARTS; ++q) {
            const uint32_t a = q < world ? s_total[2 * q] : 0u, b = q < world ? s_total[2 * q + 1] : 0u;
            tot->n_nodes[q] = a;
            tot->n_src[q] = b;
            tot->off_nodes[q] = on;
            tot->off_src[q] = os;
            on += a;
            os += b;
        }
    }
}

// Index of every sent node inside the tree of its destination (breadth-first order is kept, so the
// children of an open node stay contiguous) and first particle slot of every sent leaf.
// letidx / pidx: [destination][node].
__global__ void __launch_bounds__(256) let_index_kernel(const NodeRec *__restrict__ nodes, uint32_t n_nodes,
                                                        const uint16_t *__restrict__ open,
                                                        const uint16_t *__restrict__ reach, int world, int rank,
                                                        const uint32_t *__restrict__ tile_base, uint32_t tiles_pad,
                                                        uint32_t *__restrict__ letidx, uint32_t *__restrict__ pidx) {
    constexpr int PER = LET_TILE / 256;
    __shared__ uint32_t s_warp[2][8];
    const uint32_t x0 = blockIdx.x * LET_TILE + threadIdx.x * PER;
    uint32_t r[PER], sp[PER], cnt[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const uint32_t x = x0 + u;
        r[u] = sp[u] = cnt[u] = 0;
        if (x < n_nodes) {
            r[u] = reach[x];
            if (r[u]) {
                const NodeRec nd = nodes[x];
                cnt[u] = nd.count;
                sp[u] = (nd.nchild_level & 0xffu) == 0 ? (r[u] & open[x]) : 0u;
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        uint32_t a = 0, b = 0;  // this thread's nodes / particles for q
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            a += (r[u] >> q) & 1u;
            b += ((sp[u] >> q) & 1u) ? cnt[u] : 0u;
        }
        uint32_t ia = a, ib = b;  // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) {
                ia += ua;
                ib += ub;
            }
        }
        __syncthreads();
        if (lane == 31) {
            s_warp[0][warp] = ia;
            s_warp[1][warp] = ib;
        }
        __syncthreads();
        uint32_t base_a = tile_base[(size_t)(2 * q) * tiles_pad + blockIdx.x];
        uint32_t base_b = tile_base[(size_t)(2 * q + 1) * tiles_pad + blockIdx.x];
        for (int w = 0; w < warp; ++w) {
            base_a += s_warp[0][w];
            base_b += s_warp[1][w];
        }
        uint32_t ea = base_a + ia - a, eb = base_b + ib - b;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const uint32_t x = x0 + u;
            if ((r[u] >> q) & 1u) {
                letidx[(size_t)q * n_nodes + x] = ea++;
                if ((sp[u] >> q) & 1u) {
                    pidx[(size_t)q * n_nodes + x] = eb;
                    eb += cnt[u];
                }
            }
        }
    }
}

constexpr uint32_t LET_NO_NODE = 0xffffffffu;

// The records (and the particles of open leaves) into the send buffers; the map of the boundary nodes.
__global__ void __launch_bounds__(256) let_emit_kernel(const NodeRec *__restrict__ nodes, uint32_t n_nodes,
                                                       const float4 *__restrict__ sorted,
                                                       const uint16_t *__restrict__ open,
                                                       const uint16_t *__restrict__ reach, int world, int rank,
                                                       const uint32_t *__restrict__ letidx,
                                                       const uint32_t *__restrict__ pidx,
                                                       const LetTotals *__restrict__ tot,
                                                       const BuildState *__restrict__ st,
                                                       NodeRec *__restrict__ send_nodes, float4 *__restrict__ send_src,
                                                       uint32_t *__restrict__ bmap) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const uint32_t r = reach[x];
    if (!r) return;
    const NodeRec nd = nodes[x];
    const uint32_t nc = nd.nchild_level & 0xffu, level = nd.nchild_level >> 8, op = open[x];
    const bool first = x == st->level_begin[level], last = x + 1 == st->level_begin[level + 1];
    for (int q = 0; q < world; ++q) {
        if (q == rank || !((r >> q) & 1u)) continue;
        const uint32_t me = letidx[(size_t)q * n_nodes + x];
        NodeRec o;
        o.cm = nd.cm;
        o.first_child = 0;
        o.nchild_level = level << 8;
        o.begin = 0;
        o.count = 0;  // no children, no particles: a pruned node, accepted by the walk as it is
        if ((op >> q) & 1u) {
            if (nc) {
                o.first_child = letidx[(size_t)q * n_nodes + nd.first_child];
                o.nchild_level = nd.nchild_level;
                o.count = nd.count;
            } else {
                const uint32_t p0 = pidx[(size_t)q * n_nodes + x];
                o.begin = p0;
                o.count = nd.count;
                float4 *dst = send_src + tot->off_src[q] + p0;
                for (uint32_t i = 0; i < nd.count; ++i) dst[i] = sorted[nd.begin + i];
            }
        }
        send_nodes[tot->off_nodes[q] + me] = o;
        if (first) bmap[((size_t)q * TOP_LEVELS + level) * 2 + 0] = me;
        if (last) bmap[((size_t)q * TOP_LEVELS + level) * 2 + 1] = me;
    }
}

// Received records: links and particle ranges are relative to the sender's block.
__global__ void __launch_bounds__(256) let_rebase_kernel(NodeRec *__restrict__ nodes, uint32_t n, uint32_t node_base,
                                                         uint32_t src_base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    NodeRec r = nodes[i];
    if (r.nchild_level & 0xffu) r.first_child += node_base;
    else if (r.count) r.begin += src_base;
    nodes[i] = r;
}

// collect_boundary for trees that are not stored with their own numbering: gi[part][level][side] is the
// index of the boundary node in the joined array (LET_NO_NODE: the part has no such level).
__global__ void __launch_bounds__(2 * TOP_LEVELS * 9) let_collect_boundary(const NodeRec *__restrict__ nodes,
                                                                           const uint32_t *__restrict__ gi,
                                                                           BoundaryRec *__restrict__ out) {
    const int q = blockIdx.x;
    const int t = threadIdx.x / 9, j = threadIdx.x % 9;
    const uint32_t g = gi[(size_t)q * TOP_LEVELS * 2 + t];
    if (g == LET_NO_NODE) return;
    const NodeRec nd = nodes[g];
    BoundaryRec *o = out + (size_t)q * TOP_LEVELS * 2 + t;
    if (j == 0) o->node = nd;
    else if (j - 1 < (int)(nd.nchild_level & 0xffu)) o->child[j - 1] = nodes[nd.first_child + j - 1];
}

// Index of every boundary node in the joined array: what the sender's map says (its tree arrived pruned
// and renumbered), the level table for this rank's own tree, LET_NO_NODE where the part has no such level
// or the node was left out (it then lies below a pruned node: a complete cell; the chain of partial cells
// is always sent).
__global__ void let_gi_kernel(const uint32_t *__restrict__ bmap_recv, const PartPack *__restrict__ packs,
                              PartBases bases, int rank, bool without_own, uint32_t *__restrict__ gi) {
    const int p = blockIdx.x, t = threadIdx.x, l = t >> 1, side = t & 1;
    const PartPack &pk = packs[p];
    uint32_t g = LET_NO_NODE;
    if (pk.n_nodes && l < (int)pk.n_levels && !(without_own && p == rank)) {
        if (p == rank) g = bases.node_base[p] + (side ? pk.level_begin[l + 1] - 1 : pk.level_begin[l]);
        else {
            const uint32_t m = bmap_recv[(size_t)p * TOP_LEVELS * 2 + t];
            if (m != LET_NO_NODE) g = bases.node_base[p] + m;
        }
    }
    gi[(size_t)p * TOP_LEVELS * 2 + t] = g;
}

// Two-phase walk: the cells that this rank's own tree holds a share of are the first / last cells of its
// levels; where another rank's tree has the same cell, that node is marked NODE_SHARE (the walk over the
// others' trees then accepts it only where the whole cell would be accepted: a share's centre of mass lies
// further from this rank's particles than the cell's).
__global__ void let_mark_shares_kernel(const uint32_t *__restrict__ gi, const PartPack *__restrict__ packs, int rank,
                                       NodeRec *__restrict__ nodes) {
    const int p = blockIdx.x, t = threadIdx.x, l = t >> 1, side = t & 1;
    const PartPack &pk = packs[p], &own = packs[rank];
    if (p == rank || pk.n_nodes == 0 || own.n_nodes == 0 || l >= (int)pk.n_levels || l >= (int)own.n_levels) return;
    const uint32_t g = gi[(size_t)p * TOP_LEVELS * 2 + t];
    if (g == LET_NO_NODE) return;
    const uint64_t pre = pk.prefix[l][side];
    if (pre == own.prefix[l][0] || pre == own.prefix[l][1]) nodes[g].nchild_level |= NODE_SHARE;
}

__global__ void let_cuts_kernel(const uint64_t *__restrict__ keys, uint32_t n, const uint64_t *__restrict__ split,
                                int world, uint32_t *__restrict__ cuts, uint32_t *__restrict__ cnt_row) {
    const int q = threadIdx.x;  // cuts[q] = first sorted key >= split[q]; cuts[world] = n
    if (q > world) return;
    uint32_t lo = 0, hi = n;
    if (q == world) lo = n;
    else
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (keys[mid] < split[q]) lo = mid + 1;  // split has no low bits: decided by the sorted leading bits
            else hi = mid;
        }
    cuts[q] = q == 0 ? 0u : lo;
    __syncthreads();
    if (q < world) cnt_row[q] = cuts[q + 1] - cuts[q];
}

// Splitters = the world-quantiles of the union of the ranks' samples.  Every rank's sample is sorted (it
// was drawn from its sorted keys), so the position of an element in the sorted union is the sum of its
// lower bounds in the other lists (ties: lower list first, then position): no sort, one launch.
__global__ void __launch_bounds__(256) let_splitters_kernel(const uint64_t *__restrict__ samples, int world,
                                                            uint64_t *__restrict__ split) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, m = world * LET_SAMPLE;
    if (e == 0) {
        split[0] = 0ull;
        split[world] = ~0ull;
    }
    if (e >= m) return;
    const int list = e / LET_SAMPLE, pos = e % LET_SAMPLE;
    const uint64_t key = samples[e];
    int rank_in_union = pos;
    for (int l = 0; l < world; ++l) {
        if (l == list) continue;
        const uint64_t *a = samples + (size_t)l * LET_SAMPLE;
        int lo = 0, hi = LET_SAMPLE;  // elements of list l before this one: < key, or == key in a lower list
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const bool before = l < list ? a[mid] <= key : a[mid] < key;
            if (before) lo = mid + 1;
            else hi = mid;
        }
        rank_in_union += lo;
    }
    for (int q = 1; q < world; ++q)
        if (rank_in_union == (int)((size_t)q * m / world)) split[q] = key;
}

__global__ void __launch_bounds__(256) let_sample_kernel(const uint64_t *__restrict__ keys, uint32_t n, int m,
                                                         uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    // (the low bits are cleared: the local keys are sorted by their leading LET_SPLIT_BITS bits only)
    if (j < m) out[j] = n ? keys[(size_t)j * n / m] & ~((1ull << (63 - LET_SPLIT_BITS)) - 1ull) : ~0ull;
}

// One multi-GPU Barnes-Hut step with locally essential trees.  d_local: this rank's block [lo, hi) of
// the records; d_out: its accelerations.  Collective: every rank of the communicator must call it.
static int sharded_let_dev(pcuda_ctx *ctx, int world, int rank, size_t n_total, size_t lo, size_t hi,
                           float theta, float eps, const float *d_local, float *d_out) {
    cudaStream_t st = ctx->stream;
    pcuda_forest *f = nullptr;
    PCUDA_TRY(forest_of(ctx, &f));
    pcuda_tree *t = f->local;
    const size_t n_local = hi - lo;
    // tuning hook bh_let_trace: wall-clock time of every stage (the stream is synchronised at each mark)
    auto t_prev = std::chrono::steady_clock::now();
    std::string trace;
    auto mark = [&](const char *what) {
        if (!g_let_trace) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        char buf[64];
        snprintf(buf, sizeof buf, " %s %.3f", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        trace += buf;
        t_prev = now;
    };
    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    if (!f->h_let)
        PCUDA_CUDA_TRY(ctx, cudaHostAlloc((void **)&f->h_let,
                                          (2 * MAX_PARTS * MAX_PARTS + MAX_PARTS * TOP_LEVELS * 2) * sizeof(uint32_t) +
                                              MAX_PARTS * sizeof(LetTotals),
                                          cudaHostAllocDefault));
    uint32_t *h_cnt = f->h_let;                                   // world x MAX_PARTS particle counts
    LetTotals *h_tot = reinterpret_cast<LetTotals *>(f->h_let + 2 * MAX_PARTS * MAX_PARTS);  // per sender
    uint32_t *h_gi = reinterpret_cast<uint32_t *>(h_tot + MAX_PARTS);  // world x TOP_LEVELS x 2

    // ---- A: particles to the owners of their key ranges ----------------------------------------------
    phase_begin(ctx, PH_COMM);
    tree_reset<3>(ctx, t, 0);
    PCUDA_CUDA_TRY(ctx, f->let_boxes.ensure(8 * sizeof(float)));
    PCUDA_CUDA_TRY(ctx, f->let_box_all.ensure((size_t)world * 8 * sizeof(float)));
    PCUDA_TRY(local_box(ctx, t, d_local, n_local, f->let_boxes.as<float>()));
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, f->let_boxes.p, f->let_box_all.p, 8 * sizeof(float)));
    PCUDA_TRY(frame_from_boxes(ctx, t, f->let_box_all.as<float>(), world, n_total));
    mark("frame");
    const Frame *d_frame = t->d_frame.as<Frame>();
    const size_t nl1 = std::max<size_t>(n_local, 1);
    for (int i = 0; i < 2; ++i) {
        PCUDA_CUDA_TRY(ctx, f->let_keys[i].ensure(nl1 * sizeof(uint64_t)));
        PCUDA_CUDA_TRY(ctx, f->let_idx[i].ensure(nl1 * sizeof(uint32_t)));
        PCUDA_CUDA_TRY(ctx, f->sample[i].ensure((size_t)world * LET_SAMPLE * sizeof(uint64_t)));
    }
    PCUDA_CUDA_TRY(ctx, f->split.ensure((MAX_PARTS + 1) * sizeof(uint64_t)));
    PCUDA_CUDA_TRY(ctx, f->counts.ensure((MAX_PARTS + 1) * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->let_cuts.ensure((MAX_PARTS + 2) * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->let_cnt_mat.ensure((size_t)MAX_PARTS * MAX_PARTS * sizeof(uint32_t)));
    int cur = 0;
    size_t tmp = 0;
    if (n_local) {  // keys of the local block, sorted (stable: equal keys keep global-index order)
        launch_encode<3>(ctx, d_local, 4, n_local, d_frame, f->let_keys[0].as<uint64_t>(), f->let_idx[0].as<uint32_t>());
        cub::DoubleBuffer<uint64_t> kb(f->let_keys[0].as<uint64_t>(), f->let_keys[1].as<uint64_t>());
        cub::DoubleBuffer<uint32_t> vb(f->let_idx[0].as<uint32_t>(), f->let_idx[1].as<uint32_t>());
        // only the top LET_SPLIT_BITS bits are sorted (4 radix passes instead of 8): the splitters carry no
        // lower bits, so the destination of a key depends on those bits alone, and the receiver sorts fully
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int)n_local, 63 - LET_SPLIT_BITS, 63, st));
        PCUDA_CUDA_TRY(ctx, f->sel_tmp.ensure(tmp));
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(f->sel_tmp.p, tmp, kb, vb, (int)n_local, 63 - LET_SPLIT_BITS, 63, st));
        cur = kb.selector;
        ctx->launches += 6;
    }
    mark("local_sort");
    const uint64_t *lkeys = f->let_keys[cur].as<uint64_t>();
    const uint32_t *lperm = f->let_idx[cur].as<uint32_t>();
    // splitters: the world-quantiles of an all-gathered regular sample of the sorted local keys
    uint64_t *my_sample = f->sample[0].as<uint64_t>() + (size_t)rank * LET_SAMPLE;
    let_sample_kernel<<<(LET_SAMPLE + 255) / 256, 256, 0, st>>>(lkeys, (uint32_t)n_local, LET_SAMPLE, my_sample);
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, my_sample, f->sample[0].p, LET_SAMPLE * sizeof(uint64_t)));
    let_splitters_kernel<<<(world * LET_SAMPLE + 255) / 256, 256, 0, st>>>(f->sample[0].as<uint64_t>(), world,
                                                                           f->split.as<uint64_t>());
    uint32_t *d_mat = f->let_cnt_mat.as<uint32_t>();
    uint32_t *d_row = d_mat + (size_t)rank * MAX_PARTS;
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_row, 0, MAX_PARTS * sizeof(uint32_t), st));
    let_cuts_kernel<<<1, 32, 0, st>>>(lkeys, (uint32_t)n_local, f->split.as<uint64_t>(), world, f->let_cuts.as<uint32_t>(),
                                      d_row);
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 3 + 10;
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, d_row, d_mat, MAX_PARTS * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h_cnt, d_mat, (size_t)world * MAX_PARTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    // records in local key order: the send buffer (destination ranges are contiguous).  No index travels:
    // the accelerations come back in exactly this order (block by block), and lperm maps it to the rows.
    PCUDA_CUDA_TRY(ctx, f->let_send_rec.ensure(nl1 * sizeof(float4)));
    if (n_local) launch_gather<3>(ctx, d_local, 4, true, n_local, lperm, f->let_send_rec.as<float4>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));  // (1) the count matrix
    mark("splitters+counts");
    size_t send_off[MAX_PARTS], send_cnt[MAX_PARTS], recv_off[MAX_PARTS], recv_cnt[MAX_PARTS];
    size_t so[MAX_PARTS], sb_[MAX_PARTS], ro[MAX_PARTS], rb[MAX_PARTS];
    size_t n_mine = 0, s_off = 0;
    for (int q = 0; q < world; ++q) {
        send_off[q] = s_off;
        send_cnt[q] = h_cnt[(size_t)rank * MAX_PARTS + q];
        s_off += send_cnt[q];
        recv_off[q] = n_mine;
        recv_cnt[q] = h_cnt[(size_t)q * MAX_PARTS + rank];
        n_mine += recv_cnt[q];
    }
    if (s_off != n_local)
        return fail(ctx, PCUDA_ERR_NCCL, "key ranges do not cover the local block (%zu of %zu particles)", s_off, n_local);
    const size_t nm1 = std::max<size_t>(n_mine, 1);
    PCUDA_CUDA_TRY(ctx, f->let_recv_rec.ensure(nm1 * sizeof(float4)));
    for (int q = 0; q < world; ++q) {
        const size_t w = sizeof(float4);
        so[q] = send_off[q] * w, sb_[q] = send_cnt[q] * w, ro[q] = recv_off[q] * w, rb[q] = recv_cnt[q] * w;
    }
    PCUDA_TRY(nccl_alltoallv(ctx, f->let_send_rec.p, so, sb_, f->let_recv_rec.p, ro, rb));
    mark("a2a_particles");
    phase_end(ctx, PH_COMM);

    // ---- B: the tree of this rank's key range ---------------------------------------------------------
    phase_begin(ctx, PH_BUILD);
    tree_reset<3>(ctx, t, n_mine);
    PCUDA_CUDA_TRY(ctx, t->scan_in.ensure(sizeof(BuildState)));
    if (n_mine) {
        for (int i = 0; i < 2; ++i) {
            PCUDA_CUDA_TRY(ctx, t->keys[i].ensure(n_mine * sizeof(uint64_t)));
            PCUDA_CUDA_TRY(ctx, t->perm[i].ensure(n_mine * sizeof(uint32_t)));
        }
        const float *recv = f->let_recv_rec.as<float>();
        launch_encode<3>(ctx, recv, 4, n_mine, d_frame, t->keys[0].as<uint64_t>(), t->perm[0].as<uint32_t>());
        cub::DoubleBuffer<uint64_t> kb(t->keys[0].as<uint64_t>(), t->keys[1].as<uint64_t>());
        cub::DoubleBuffer<uint32_t> vb(t->perm[0].as<uint32_t>(), t->perm[1].as<uint32_t>());
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, kb, vb, (int)n_mine, 0, 63, st));
        PCUDA_CUDA_TRY(ctx, t->cub_tmp.ensure(tmp));
        PCUDA_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(t->cub_tmp.p, tmp, kb, vb, (int)n_mine, 0, 63, st));
        t->cur = kb.selector;
        PCUDA_CUDA_TRY(ctx, t->sorted.ensure(n_mine * sizeof(float4)));
        launch_gather<3>(ctx, recv, 4, true, n_mine, t->d_perm(), t->sorted.as<float4>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 10;
        mark("sort_received");
        PCUDA_TRY(build_levels<3>(ctx, t, n_mine));  // (2) synchronises: level table
        mark("tree");
    }
    // Two-phase walk (from 8 ranks on; tuning hook bh_let_overlap = 0 / 1: never / always).  The interactions of this rank's
    // particles with this rank's own tree need nothing from the other ranks: that walk starts now, on a
    // second stream, and leaves the last SMs to the pruning kernels and the exchanges of stages C and D,
    // which run beside it.  What arrives is walked afterwards (stage E) and added to the same rows.
    // A cell that straddles this rank's key range and another's is then seen as two shares (this rank's: the
    // first / last node of a level of its tree; the others': let_mark_shares_kernel), and a share is accepted
    // only where the whole cell would be (traverse2_kernel), so the walk is nowhere coarser than the one walk.
    const bool overlap = (g_let_overlap < 0 ? world >= 8 : g_let_overlap != 0) && n_mine > 0;
    struct Join {  // no return path leaves the second stream running behind the first
        pcuda_forest *f;
        cudaStream_t st;
        bool armed = false;
        ~Join() {
            if (armed) cudaStreamWaitEvent(st, f->ev_join, 0);
        }
    } join{f, st};
    PCUDA_CUDA_TRY(ctx, f->route_acc_send.ensure(std::max<size_t>(n_mine, 1) * 3 * sizeof(float)));
    if (overlap) {
        if (!f->walk_stream) {
            PCUDA_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&f->walk_stream, cudaStreamNonBlocking));
            PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&f->ev_fork, cudaEventDisableTiming));
            PCUDA_CUDA_TRY(ctx, cudaEventCreateWithFlags(&f->ev_join, cudaEventDisableTiming));
        }
        PCUDA_CUDA_TRY(ctx, f->let_root0.ensure(2 * sizeof(uint32_t)));  // [0]: the root (node 0); [1]: stop flag
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(f->let_root0.p, 0, 2 * sizeof(uint32_t), st));
        PCUDA_CUDA_TRY(ctx, cudaEventRecord(f->ev_fork, st));
        PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(f->walk_stream, f->ev_fork, 0));
        ForestView own{};
        own.nodes = t->nodes.as<NodeRec>();
        own.src = t->sorted.as<float4>();
        own.d_roots = f->let_root0.as<uint32_t>();
        own.n_roots = 1;
        own.d_level_begin = t->scan_in.as<BuildState>()->level_begin;
        own.stream = f->walk_stream;
        own.reserve_sms = (unsigned)g_let_reserve;
        if (g_let_stop) own.d_stop = f->let_root0.as<uint32_t>() + 1;
        const int rc = traverse_sorted(ctx, t, t->sorted.as<float4>(), t->d_keys(), t->d_perm(), n_mine, theta, eps,
                                       f->route_acc_send.as<float>(), nullptr, &own);
        cudaEventRecord(f->ev_join, f->walk_stream);
        join.armed = true;
        PCUDA_TRY(rc);
    }
    PartPack *d_packs = f->packs.as<PartPack>();
    fill_pack<<<1, 64, 0, st>>>(t->nodes.as<NodeRec>(), t->moments.as<double>(), n_mine ? t->d_keys() : nullptr,
                                t->scan_in.as<BuildState>(), (uint32_t)t->n_nodes, (uint32_t)t->n_levels, d_packs + rank);
    // where this rank's targets are: the top of the tree, refined while a cell is heavy (LetDomain)
    const uint32_t nn = (uint32_t)t->n_nodes;
    const size_t nn1 = std::max<size_t>(nn, 1);
    PCUDA_CUDA_TRY(ctx, f->let_dom.ensure(sizeof(LetDomain)));
    PCUDA_CUDA_TRY(ctx, f->let_dom_all.ensure((size_t)world * sizeof(LetDomain)));
    PCUDA_CUDA_TRY(ctx, f->let_parent.ensure(nn1 * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->let_index.ensure((size_t)2 * world * nn1 * sizeof(uint32_t)));
    if (nn) {
        const unsigned nb256 = (nn + 255) / 256;
        const uint32_t tau = (uint32_t)std::max<size_t>((size_t)g_seg_max, (8 * n_mine + LET_DOM_MAX - 1) / LET_DOM_MAX);
        uint32_t *d_flag = f->let_index.as<uint32_t>(), *d_idx = d_flag + nn1;  // scratch: free until let_index_kernel
        let_parent_kernel<<<nb256, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, f->let_parent.as<uint32_t>());
        let_dom_flag_kernel<<<nb256, 256, 0, st>>>(t->nodes.as<NodeRec>(), f->let_parent.as<uint32_t>(), nn, tau, d_flag);
        PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_flag, d_idx, (int)nn, st));
        PCUDA_CUDA_TRY(ctx, f->sel_tmp.ensure(tmp));
        PCUDA_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(f->sel_tmp.p, tmp, d_flag, d_idx, (int)nn, st));
        let_domain_kernel<<<nb256, 256, 0, st>>>(t->nodes.as<NodeRec>(), t->d_keys(), d_flag, d_idx, nn, tau, d_frame,
                                                 f->let_dom.as<LetDomain>());
        let_domain_box_kernel<<<1, 256, 0, st>>>(f->let_dom.as<LetDomain>());
        ctx->launches += 6;
    } else {
        PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(f->let_dom.p, 0, 32, st));  // n = 0: no targets here
    }
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 1;
    mark("pack+domain");
    phase_end(ctx, PH_BUILD);

    // ---- C: locally essential trees ---------------------------------------------------------------------
    phase_begin(ctx, PH_COMM3);
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, d_packs + rank, d_packs, sizeof(PartPack)));
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, f->let_dom.p, f->let_dom_all.p, sizeof(LetDomain)));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_packs, d_packs, world * sizeof(PartPack), cudaMemcpyDeviceToHost, st));
    const uint32_t n_tiles = (nn + LET_TILE - 1) / LET_TILE, tiles_pad = (std::max<uint32_t>(n_tiles, 1) + 31u) & ~31u;
    PCUDA_CUDA_TRY(ctx, f->let_open.ensure(nn1 * sizeof(uint16_t)));
    PCUDA_CUDA_TRY(ctx, f->let_reach.ensure(nn1 * sizeof(uint16_t)));
    PCUDA_CUDA_TRY(ctx, f->let_tile_cnt.ensure((size_t)2 * MAX_PARTS * tiles_pad * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->let_totals.ensure((size_t)MAX_PARTS * sizeof(LetTotals)));
    PCUDA_CUDA_TRY(ctx, f->let_bmap_send.ensure((size_t)MAX_PARTS * TOP_LEVELS * 2 * sizeof(uint32_t)));
    PCUDA_CUDA_TRY(ctx, f->let_bmap_recv.ensure((size_t)MAX_PARTS * TOP_LEVELS * 2 * sizeof(uint32_t)));
    // worst case: every node and every particle goes to every other rank
    PCUDA_CUDA_TRY(ctx, f->let_send_nodes.ensure(std::max<size_t>(1, (size_t)(world - 1) * nn) * sizeof(NodeRec)));
    PCUDA_CUDA_TRY(ctx, f->let_send_src.ensure(std::max<size_t>(1, (size_t)(world - 1) * n_mine) * sizeof(float4)));
    LetTotals *d_tot_all = f->let_totals.as<LetTotals>();
    LetTotals *d_tot = d_tot_all + rank;
    uint32_t *letidx = f->let_index.as<uint32_t>(), *pidx = letidx + (size_t)world * nn1;
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(f->let_bmap_send.p, 0xff, (size_t)MAX_PARTS * TOP_LEVELS * 2 * sizeof(uint32_t), st));
    PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(d_tot, 0, sizeof(LetTotals), st));
    if (nn) {
        const unsigned nb256 = (nn + 255) / 256;
        const BuildState *d_state = t->scan_in.as<BuildState>();
        let_open_kernel<<<nb256, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, d_state, f->let_dom_all.as<LetDomain>(), world,
                                               rank, theta * theta, d_frame, f->let_open.as<uint16_t>());
        const uint32_t all = ((1u << world) - 1u) & ~(1u << rank);
        let_reach_kernel<<<nb256, 256, 0, st>>>(f->let_open.as<uint16_t>(), f->let_parent.as<uint32_t>(), nn, all,
                                                f->let_reach.as<uint16_t>());
        let_count_kernel<<<n_tiles, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, f->let_open.as<uint16_t>(),
                                                  f->let_reach.as<uint16_t>(), world, f->let_tile_cnt.as<uint32_t>(),
                                                  tiles_pad);
        let_scan_kernel<<<1, 1024, 0, st>>>(f->let_tile_cnt.as<uint32_t>(), n_tiles, tiles_pad, world, d_tot);
        let_index_kernel<<<n_tiles, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, f->let_open.as<uint16_t>(),
                                                  f->let_reach.as<uint16_t>(), world, rank, f->let_tile_cnt.as<uint32_t>(),
                                                  tiles_pad, letidx, pidx);
        let_emit_kernel<<<nb256, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, t->sorted.as<float4>(),
                                               f->let_open.as<uint16_t>(), f->let_reach.as<uint16_t>(), world, rank, letidx,
                                               pidx, d_tot, d_state, f->let_send_nodes.as<NodeRec>(),
                                               f->let_send_src.as<float4>(), f->let_bmap_send.as<uint32_t>());
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 6;
    }
    mark("let_kernels");
    PCUDA_TRY(pcuda_comm_allgather_dev(ctx, d_tot, d_tot_all, sizeof(LetTotals)));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h_tot, d_tot_all, (size_t)world * sizeof(LetTotals), cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));  // (3) sizes of the trees on their way
    mark("let_counts");
    // joined arrays: block p = what rank p sent (p == rank: the whole local tree), then the top tree
    uint32_t node_base[MAX_PARTS], src_base[MAX_PARTS];
    size_t n_nodes_all = 0, n_src_all = 0;
    for (int p = 0; p < world; ++p) {
        node_base[p] = (uint32_t)n_nodes_all;
        src_base[p] = (uint32_t)n_src_all;
        n_nodes_all += p == rank ? (overlap ? 0 : nn) : h_tot[p].n_nodes[rank];
        n_src_all += p == rank ? (overlap ? 0 : n_mine) : h_tot[p].n_src[rank];
    }
    if (n_nodes_all + TOP_CAP > 0xfffffff0ull || n_src_all > 0xfffffff0ull)
        return fail(ctx, PCUDA_ERR_TREE_OVERFLOW, "joined tree does not fit 32-bit indices");
    PCUDA_CUDA_TRY(ctx, f->nodes.ensure((n_nodes_all + TOP_CAP) * sizeof(NodeRec)));
    PCUDA_CUDA_TRY(ctx, f->sorted.ensure(std::max<size_t>(n_src_all, 1) * sizeof(float4)));
    PCUDA_TRY(nccl_group_begin(ctx));       // one fused exchange
    for (int pass = 0; pass < 3; ++pass) {  // nodes, particles, boundary maps
        const size_t w = pass == 0 ? sizeof(NodeRec) : pass == 1 ? sizeof(float4) : TOP_LEVELS * 2 * sizeof(uint32_t);
        for (int q = 0; q < world; ++q) {
            if (pass == 2) {
                so[q] = (size_t)q * w, ro[q] = (size_t)q * w;
                sb_[q] = rb[q] = q == rank ? 0 : w;
            } else {
                const uint32_t *soff = pass == 0 ? h_tot[rank].off_nodes : h_tot[rank].off_src;
                const uint32_t *scnt = pass == 0 ? h_tot[rank].n_nodes : h_tot[rank].n_src;
                so[q] = (size_t)soff[q] * w;
                sb_[q] = q == rank ? 0 : (size_t)scnt[q] * w;
                ro[q] = (size_t)(pass == 0 ? node_base[q] : src_base[q]) * w;
                rb[q] = q == rank ? 0 : (size_t)(pass == 0 ? h_tot[q].n_nodes[rank] : h_tot[q].n_src[rank]) * w;
            }
        }
        const void *sp = pass == 0 ? f->let_send_nodes.p : pass == 1 ? f->let_send_src.p : f->let_bmap_send.p;
        void *rp = pass == 0 ? f->nodes.p : pass == 1 ? f->sorted.p : f->let_bmap_recv.p;
        PCUDA_TRY(nccl_alltoallv(ctx, sp, so, sb_, rp, ro, rb));
    }
    PCUDA_TRY(nccl_group_end(ctx));
    NodeRec *fn = f->nodes.as<NodeRec>();
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        const uint32_t cnt = h_tot[p].n_nodes[rank];
        if (cnt) let_rebase_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(fn + node_base[p], cnt, node_base[p], src_base[p]);
    }
    if (nn && !overlap) {
        copy_rebase_nodes<<<(nn + 255) / 256, 256, 0, st>>>(t->nodes.as<NodeRec>(), nn, node_base[rank], src_base[rank],
                                                            fn + node_base[rank]);
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->sorted.as<float4>() + src_base[rank], t->sorted.p, n_mine * sizeof(float4),
                                            cudaMemcpyDeviceToDevice, st));
    }
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += world;

    mark("a2a_let+rebase");
    // ---- D: the top tree ----------------------------------------------------------------------------------
    PartBases bases{};
    for (int p = 0; p < world; ++p) bases.node_base[p] = node_base[p];
    PCUDA_CUDA_TRY(ctx, f->let_gi.ensure((size_t)MAX_PARTS * TOP_LEVELS * 2 * sizeof(uint32_t)));
    let_gi_kernel<<<world, 2 * TOP_LEVELS, 0, st>>>(f->let_bmap_recv.as<uint32_t>(), d_packs, bases, rank, overlap,
                                                    f->let_gi.as<uint32_t>());
    if (overlap) let_mark_shares_kernel<<<world, 2 * TOP_LEVELS, 0, st>>>(f->let_gi.as<uint32_t>(), d_packs, rank, fn);
    let_collect_boundary<<<world, 2 * TOP_LEVELS * 9, 0, st>>>(fn, f->let_gi.as<uint32_t>(), f->stage.as<BoundaryRec>());
    PCUDA_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(h_gi, f->let_gi.p, (size_t)world * TOP_LEVELS * 2 * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->h_stage, f->stage.p, (size_t)world * TOP_LEVELS * 2 * sizeof(BoundaryRec),
                                        cudaMemcpyDeviceToHost, st));
    PCUDA_CUDA_TRY(ctx, cudaStreamSynchronize(st));  // (4) boundary records and their indices
    std::vector<NodeRec> top;
    std::vector<uint32_t> roots;
    if (overlap) f->h_packs[rank].n_nodes = 0;  // the top tree of the OTHER ranks' trees
    PCUDA_TRY(merge_top_tree(ctx, world, f->h_packs, f->h_stage, node_base, (uint32_t)n_nodes_all, top, roots, h_gi));
    if (!top.empty())
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(fn + n_nodes_all, top.data(), top.size() * sizeof(NodeRec),
                                            cudaMemcpyHostToDevice, st));
    if (!roots.empty())
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(f->roots.p, roots.data(), roots.size() * sizeof(uint32_t),
                                            cudaMemcpyHostToDevice, st));
    ForestView fv{};
    fv.nodes = fn;
    fv.src = f->sorted.as<float4>();
    fv.d_roots = f->roots.as<uint32_t>();
    fv.n_roots = (uint32_t)roots.size();

    mark("top_tree");
    // ---- E: walk, accelerations back along the path the particles came ------------------------------------
    // The received records lie block by block in the order their owners sent them, and t->d_perm() maps a
    // sorted particle to its place there: the walk writes every row straight into that place, the blocks go
    // back with the counts of step A swapped, and the owner scatters them through its own sort permutation.
    PCUDA_CUDA_TRY(ctx, f->route_acc_recv.ensure(nl1 * 3 * sizeof(float)));
    phase_end(ctx, PH_COMM3);
    phase_begin(ctx, PH_COMPUTE);  // (two-phase walk: what is left of the first phase, and the second)
    if (overlap) {
        // tuning hook bh_let_stop: the others' trees are here, the first phase stops taking groups (it ran
        // without the reserved SMs) and what it left is walked on the whole GPU in front of the second phase.
        // Measured on 4 B200s: 8.51 ms against 8.40 without — a walk ends with a tail of about half a group
        // time (a group keeps a warp busy for ~0.7 ms), and this makes three tails out of two.
        if (g_let_stop) PCUDA_CUDA_TRY(ctx, cudaMemsetAsync(f->let_root0.as<uint32_t>() + 1, 1, sizeof(uint32_t), st));
        PCUDA_CUDA_TRY(ctx, cudaStreamWaitEvent(st, f->ev_join, 0));
        join.armed = false;
        if (g_let_stop) {
            ForestView rest{};
            rest.nodes = t->nodes.as<NodeRec>();
            rest.src = t->sorted.as<float4>();
            rest.d_roots = f->let_root0.as<uint32_t>();
            rest.n_roots = 1;
            rest.d_level_begin = t->scan_in.as<BuildState>()->level_begin;
            rest.continue_groups = true;
            PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>(), t->d_keys(), t->d_perm(), n_mine, theta, eps,
                                      f->route_acc_send.as<float>(), nullptr, &rest));
        }
        fv.accumulate = true;
        fv.reuse_groups = true;
    }
    if (n_mine && (!overlap || fv.n_roots))
        PCUDA_TRY(traverse_sorted(ctx, t, t->sorted.as<float4>(), t->d_keys(), t->d_perm(), n_mine, theta, eps,
                                  f->route_acc_send.as<float>(), nullptr, &fv));
    mark("walk");
    phase_end(ctx, PH_COMPUTE);
    phase_begin(ctx, PH_COMM2);
    for (int q = 0; q < world; ++q) {
        const size_t w = 3 * sizeof(float);
        so[q] = recv_off[q] * w, sb_[q] = recv_cnt[q] * w, ro[q] = send_off[q] * w, rb[q] = send_cnt[q] * w;
    }
    PCUDA_TRY(nccl_alltoallv(ctx, f->route_acc_send.p, so, sb_, f->route_acc_recv.p, ro, rb));
    if (n_local) {
        scatter_rows<<<(unsigned)((n_local + 255) / 256), 256, 0, st>>>(f->route_acc_recv.as<float>(), lperm, (int)n_local,
                                                                        0u, d_out);
        PCUDA_CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    mark("route_back");
    if (g_let_trace)
        fprintf(stderr, "[let rank %d/%d n_mine %zu nodes %u let_in nodes %zu src %zu]%s\n", rank, world, n_mine, nn,
                n_nodes_all - nn, n_src_all - n_mine, trace.c_str());
    phase_end(ctx, PH_COMM2);
    return PCUDA_OK;
}

static int sharded_dev_impl(pcuda_ctx *ctx, const float *d_local, size_t n_local, size_t n_total,
                            float theta, float eps, float *d_gathered, float *d_out) {
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    const size_t lo = std::min(n_total, (size_t)rank * cap), hi = std::min(n_total, lo + cap);
    if (n_local != hi - lo)
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "rank %d of %d must own %zu of %zu particles (contiguous blocks of %zu), got %zu",
                    rank, world, hi - lo, n_total, cap, n_local);
    const int how = g_forest ? g_forest : ctx->bh_build;  // 0 = automatic
    const bool forest_ok = world > 1 && world <= MAX_PARTS && ctx->order == 1 && g_tpl == 2 && !g_variant;
    // locally essential trees: nothing is replicated.  Measured at N = 10M (r02, ms per step, LET / partitioned /
    // replicated): 2 GPUs 14.97 / 14.93 / 14.86, 4 GPUs 8.18 / 8.34 / 8.93, 8 GPUs 5.01 / 5.31 / 6.04 — hence
    // from 3 GPUs on, when every rank has a range worth a tree
    if (forest_ok && (how == 3 || (how == 0 && world >= 3 && n_total >= (size_t)world * 65536)))
        return sharded_let_dev(ctx, world, rank, n_total, lo, hi, theta, eps, d_local, d_out);
    float *slot = d_gathered + (size_t)rank * cap * 4;
    phase_begin(ctx, PH_COMM);
    if (n_local)
        PCUDA_CUDA_TRY(ctx, cudaMemcpyAsync(slot, d_local, n_local * 16, cudaMemcpyDeviceToDevice,
                                            ctx->stream));
    if (world > 1) PCUDA_TRY(pcuda_comm_allgather_dev(ctx, slot, d_gathered, cap * 16));
    phase_end(ctx, PH_COMM);
    const bool forest = how == 1 || (how == 0 && world >= 4);
    if (forest && forest_ok && n_total >= (size_t)world)
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



// Arguments are validated before the first collective; a failure after it (out of memory, an
// inconsistent exchange) poisons the communicator (comm.cu: nccl_poison) so that this rank fails fast from
// then on instead of entering collectives its peers have left.
int sharded_dev(pcuda_ctx *ctx, const float *d_local, size_t n_local, size_t n_total, float theta, float eps,
                float *d_gathered, float *d_out) {
    if (nccl_poisoned(ctx))
        return fail(ctx, PCUDA_ERR_NCCL, "the communicator is unusable: an earlier multi-GPU step failed on this rank");
    int world = 1, rank = 0;
    nccl_world(ctx, &world, &rank);
    const size_t cap = std::max<size_t>(1, (n_total + world - 1) / world);
    const size_t lo = std::min(n_total, (size_t)rank * cap), hi = std::min(n_total, lo + cap);
    if (n_local != hi - lo)  // before any collective: the other ranks are not left waiting
        return fail(ctx, PCUDA_ERR_INVALID_ARGUMENT,
                    "rank %d of %d must own %zu of %zu particles (contiguous blocks of %zu), got %zu", rank, world,
                    hi - lo, n_total, cap, n_local);
    const int s = sharded_dev_impl(ctx, d_local, n_local, n_total, theta, eps, d_gathered, d_out);
    if (s != PCUDA_OK && world > 1) nccl_poison(ctx);
    return s;
}

}  // namespace bh
}  // namespace pcuda

using namespace pcuda;

extern "C" {

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

}  // extern "C"
