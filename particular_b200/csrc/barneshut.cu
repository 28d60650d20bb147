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
//
// Translation units: bh_build.cu (K2-K4 level-wise / single-block builds, quadrupoles, f64 layer),
// bh_radix_build.cu (K4 in one pass), bh_traverse.cu (K5), bh_multigpu.cu (sharded paths); this file
// holds the one-shot drivers and the C ABI entry points.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "bh.cuh"

namespace pcuda {
namespace bh {

// Tuning / test hooks (pcuda_debug_set); declared in bh.cuh.
int g_level_build = 0;            // 1 = level-wise build instead of the one-pass build, 2 = one-pass build at
                                  // every size (also where build_small would run)
uint32_t g_small_level = 131072;  // level-wise build: levels up to this many nodes take the node-x-digit path
int g_variant = 0;                // experimental traversal variants
bool g_count = true;              // instrumentation of the traversal (pcuda_tree_last_counters)
int g_seg_max = 256;              // largest cell (in targets) that is cut into groups
int g_tpl = 2;                    // targets per lane in the traversal: 1 (groups of 32) or 2 (groups of 64)
int g_route = 0;                  // accelerations to their owners: 0 = automatic, 1 = all-gather, 2 = all-to-all
int g_tree_groups = 1;            // target groups from the tree's nodes when the targets are its own particles
int g_let_overlap = -1;           // locally essential trees: walk the own tree while the others' trees travel
                                  // (-1: from 8 ranks on, where it was measured to pay; 0: never; 1: always)
int g_let_reserve = 16;           // ... the SMs that walk leaves to the kernels beside it
int g_let_stop = 0;               // ... and whether it stops when the others' trees are here (measured: no gain)
int g_let_trace = 0;              // locally essential trees: print the wall-clock time of every stage
int g_forest = 0;                 // multi-GPU build: 0 = as the context says, 1 = partitioned, 2 = replicated,
                                  // 3 = locally essential trees

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
    if (k == "bh_variant" && value >= 0 && value <= 4) {
        bh::g_variant = value;
        return PCUDA_OK;
    }
    if (k == "bh_count") {
        bh::g_count = value != 0;
        return PCUDA_OK;
    }
    if (k == "bh_forest" && value >= 0 && value <= 3) {
        bh::g_forest = value;
        return PCUDA_OK;
    }
    if (k == "bh_tree_groups") {
        bh::g_tree_groups = value;
        return PCUDA_OK;
    }
    if (k == "bh_let_overlap" && value >= -1 && value <= 1) {
        bh::g_let_overlap = value;
        return PCUDA_OK;
    }
    if (k == "bh_let_stop") {
        bh::g_let_stop = value != 0;
        return PCUDA_OK;
    }
    if (k == "bh_let_reserve" && value >= 0 && value <= 64) {
        bh::g_let_reserve = value;
        return PCUDA_OK;
    }
    if (k == "bh_let_trace") {
        bh::g_let_trace = value;
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

