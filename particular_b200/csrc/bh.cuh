// bh.cuh — declarations shared by the Barnes-Hut translation units (barneshut.cu, bh_radix_build.cu).
// Not part of the C ABI (that is include/particular_cuda.h).
#pragma once

#include "common.cuh"

namespace pcuda {
namespace bh {

// One tree node: 32 bytes = one DRAM sector.
struct __align__(32) NodeRec {
    float4 cm;             // centre of mass {x, y, z (0 in 2-D)}, w = total mu
    uint32_t first_child;  // index of the first child (children are contiguous); 0 for leaves
    uint32_t nchild_level; // n_children | level << 8
    uint32_t begin;        // first sorted particle of the cell
    uint32_t count;        // particles in the cell
};

struct Frame {  // quantisation frame == the reference's root cube
    float origin[3];
    float ext;
    float inv;
    float mass_bound;  // n * max|mu|: bounds the |mass| of every node (not part of the tree spec)
};

template <int DIM>
struct Dims {
    static constexpr int BITS = DIM == 3 ? 21 : 31;
    static constexpr int X = 1 << DIM;
};

}  // namespace bh
}  // namespace pcuda

struct pcuda_tree {
    int dim = 3, bits = 21;
    size_t n = 0, n_nodes = 0;
    int n_levels = 0;
    uint32_t leaf_size = 16;
    double nodes_per_particle = 0.5;    // capacity guess; doubled when a build overflows
    std::vector<uint32_t> level_begin;  // n_levels + 1 entries
    pcuda::bh::Frame frame = {};
    pcuda::DevBuf keys[2], perm[2], sorted, nodes, moments, d_frame, scan_in, scan_out, cub_tmp,
        partial;
    pcuda::DevBuf rb;            // scratch of the one-pass build (bh_radix_build.cu)
    uint32_t *rb_totals = nullptr;       // level totals + ticket of rb_scan (inside `rb`), zeroed when (re)placed
    const uint32_t *d_parent = nullptr;  // parent of every node (inside `rb`); nullptr when the tree was not built
                                         // by the one-pass build (single-block / level-wise builds)
    pcuda::DevBuf quad64, quad;  // expansion order 2: traceless quadrupole per node, 6 doubles
                                 // (build) and 2 x float4 {xx, xy, xz, yy}{yz, zz, 0, 0} (traversal)
    int order = 1;
    pcuda::DevBuf sorted64;  // f64 trees: the sources in key order as double4 {x, y, z|0, mu};
                             // `moments` then holds the double-precision {com, mass} per node
    int cur = 0;  // which of keys[]/perm[] holds the sorted data
    uint64_t *d_keys() const { return keys[cur].as<uint64_t>(); }
    uint32_t *d_perm() const { return perm[cur].as<uint32_t>(); }
};

namespace pcuda {
namespace bh {

// Level table of a build in device memory (the per-level kernels and the one-pass build fill it).
struct BuildState {
    uint32_t level_begin[36];  // level l = nodes [level_begin[l], level_begin[l+1])
    uint32_t ticket[34];       // tile dispenser of each level's kernel
    uint32_t overflow;         // a level did not fit into `capacity` nodes
    uint32_t capacity;
};

constexpr int MAX_PARTS = 16;       // trees in a forest (= GPUs of a partitioned build)
constexpr int MAX_ROOTS = 736;      // start nodes of a forest walk (stack capacity of the walk - reserve - 32)
constexpr int SEG_MAX_LIMIT = 1024; // largest segment size the grouping kernels support

// Tuning / test hooks (pcuda_debug_set); defined in barneshut.cu.
extern int g_level_build;
extern uint32_t g_small_level;
extern int g_variant;
extern bool g_count;
extern int g_seg_max;
extern int g_tpl;
extern int g_route;
extern int g_forest;
extern int g_let_trace, g_let_overlap, g_let_reserve, g_let_stop;
extern int g_tree_groups;

// Double-precision layer of a traversal (tree built by build64).
struct Ext64 {
    const double4 *src64;  // sources in key order
    const double4 *cm64;   // {com, mass} per node
    const double4 *tgt64;  // targets in traversal order {x, y, z|0, _}
    double *out;
    double eps2;
};

// A forest of trees over the same root cube stored back to back (partitioned build): node and
// source arrays that replace the tree's own, and the roots the walk starts from.
// NodeRec::nchild_level, bit 31 (joined trees only): the node is one side's share of a cell whose other
// share is walked separately; the walk accepts it only where it would accept the whole cell.
constexpr uint32_t NODE_SHARE = 0x80000000u;

struct ForestView {
    const NodeRec *nodes;
    const float4 *src;
    const uint32_t *d_roots;  // device array
    uint32_t n_roots;
    // two-phase walks (locally essential trees: own tree while the others' trees are on their way)
    const uint32_t *d_n_roots = nullptr;  // number of start nodes in device memory (overrides n_roots)
    const uint32_t *d_level_begin = nullptr;  // level table (device) of the walked tree: its first / last nodes
                                          // of every level may be shares of cells (see NODE_SHARE)
    bool accumulate = false;              // add to the output rows instead of writing them
    bool reuse_groups = false;            // the target groups of the previous walk of the same targets
    bool continue_groups = false;         // ... and only those the previous walk left when it was stopped
    const uint32_t *d_stop = nullptr;     // the walk takes no more groups once this device word is non-zero
    unsigned reserve_sms = 0;             // blocks that start on the last reserve_sms SMs give them up (at most
                                          // 4 per SM do): room for the kernels of a concurrent stream
    cudaStream_t stream = nullptr;        // nullptr: the context stream
};

// ---- bh_build.cu ----
template <int DIM>
int sort_by_key(pcuda_ctx *ctx, const float *d_pos, int stride, size_t n, const Frame *d_frame,
                DevBuf keys[2], DevBuf perm[2], int *cur, DevBuf &cub_tmp);
template <int DIM>
void tree_reset(pcuda_ctx *ctx, pcuda_tree *t, size_t n);
template <int DIM>
int build_frame(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n);
template <int DIM>
int build_levels(pcuda_ctx *ctx, pcuda_tree *t, size_t n);
template <int DIM>
int build(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n, bool keys_only = false);
int build_dim(pcuda_ctx *ctx, pcuda_tree *t, uint32_t dim, const float *d_particles, size_t n,
              bool keys_only = false);
template <int DIM>
int build64(pcuda_ctx *ctx, pcuda_tree *t, const double *d_particles64, size_t n);
template <int DIM>
void launch_encode(pcuda_ctx *ctx, const float *d_pos, int stride, size_t n, const Frame *d_frame,
                   uint64_t *keys, uint32_t *idx);
template <int DIM>
void launch_gather(pcuda_ctx *ctx, const float *d_pos, int stride, bool has_mass, size_t n,
                   const uint32_t *perm, float4 *sorted);
template <int DIM>
void launch_gather64(pcuda_ctx *ctx, const double *d_pos, int stride, bool has_mass, size_t n,
                     const uint32_t *perm, double4 *sorted);
void launch_narrow(pcuda_ctx *ctx, const double *in, size_t count, float *out);
// Root cube of a cloud spread over several ranks: {lo[3], hi[3], max|mu|, 0} of the local records, and
// the frame (into t->d_frame) from the boxes of all ranks; same bits as build_frame over all particles.
int local_box(pcuda_ctx *ctx, pcuda_tree *t, const float *d_particles, size_t n, float *d_box8);
int frame_from_boxes(pcuda_ctx *ctx, pcuda_tree *t, const float *d_boxes, int world, size_t n_total);

// ---- bh_traverse.cu ----
// d_tgt == nullptr: the targets are the tree's own particles (the `&[P]` storage).
// tgt_stride: floats per target row (0 = bare positions, i.e. `dim`).
int traverse(pcuda_ctx *ctx, const pcuda_tree *t, const float *d_tgt, size_t na, float theta, float eps,
             float *d_out, int tgt_stride = 0, const double *d_tgt64 = nullptr, double *d_out64 = nullptr,
             double eps64 = 0.0);
int traverse_sorted(pcuda_ctx *ctx, const pcuda_tree *t, const float4 *tgt_sorted, const uint64_t *tgt_keys,
                    const uint32_t *tgt_perm, size_t na, float theta, float eps, float *d_out,
                    const Ext64 *x64 = nullptr, const ForestView *fv = nullptr);
int read_counters(pcuda_ctx *ctx);

// ---- bh_multigpu.cu ----
int sharded_dev(pcuda_ctx *ctx, const float *d_local, size_t n_local, size_t n_total, float theta, float eps,
                float *d_gathered, float *d_out);
int partitioned_dev(pcuda_ctx *ctx, const float *d_particles, size_t n, int parts, float theta, float eps,
                    float *d_out);

constexpr int RB_MAX_LEAF = 32;  // widest leaf window of the one-pass build (== the cap on leaf_size)

// bh_radix_build.cu: one-pass construction of the linear orthtree over t->d_keys() / t->sorted
// (n sorted particles): fills t->nodes, t->moments and *d_state; enqueue only, no synchronisation.
template <int DIM>
int radix_build_enqueue(pcuda_ctx *ctx, pcuda_tree *t, size_t n, size_t cap_nodes, BuildState *d_state);

}  // namespace bh
}  // namespace pcuda
