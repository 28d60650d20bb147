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

constexpr int RB_MAX_LEAF = 32;  // widest leaf window of the one-pass build (== the cap on leaf_size)

// bh_radix_build.cu: one-pass construction of the linear orthtree over t->d_keys() / t->sorted
// (n sorted particles): fills t->nodes, t->moments and *d_state; enqueue only, no synchronisation.
template <int DIM>
int radix_build_enqueue(pcuda_ctx *ctx, pcuda_tree *t, size_t n, size_t cap_nodes, BuildState *d_state);

}  // namespace bh
}  // namespace pcuda
