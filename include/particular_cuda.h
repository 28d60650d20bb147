/*
 * particular_cuda.h — C ABI of libparticular_cuda.so, the B200 (sm_100a) compute backend for the
 * `particular` N-body crate.
 *
 * This is the drop-in boundary: exactly what a `particular-cuda` Rust crate binds through
 * `extern "C"` (see INTEGRATION.md and rust/particular-cuda/src/ffi.rs (source only: no Rust toolchain here)) so that its
 * `cuda::BruteForce` / `cuda::BarnesHut` types can implement the crate's operator trait
 *     Interaction<Between<&[P1], &[P2]>>           (reference particular/src/lib.rs:364-370)
 * the same way the existing wgpu operator does     (reference particular/src/gpu/mod.rs:179-208).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary;
 *   - every entry point returns a pcuda_status (0 = ok, < 0 = error) and never aborts; the message
 *     of the last failure is available from pcuda_last_error();
 *   - there is NO CPU fallback: without a usable sm_100 device pcuda_create() fails;
 *   - a context is bound to one device, owns one stream and grow-only device/pinned buffers, and
 *     allows one call in flight (the Rust wrapper holds `&mut CudaContext`, as the reference holds
 *     `&mut GpuResources`, gpu/mod.rs:149-159);
 *   - wire layout follows `GravitationalField<V, S>` (#[repr(C)] {position, m},
 *     reference particular/src/gravity/mod.rs:12-18), densely packed, native endianness:
 *         affecting (sources):  f32x3: float[4]  {x,y,z,mu}    f32x2: float[3] {x,y,mu}
 *                               f64x3: double[4] {x,y,z,mu}
 *         affected  (targets):  bare positions, float[3] / float[2] / double[3];
 *                               NULL means "affected == affecting" (the `&[P]` storage,
 *                               reference storage.rs:231-241) and saves one upload;
 *         out:                  accelerations, same shape/type as the affected positions,
 *                               in affected order.
 *   - semantics of one call = reference `sequential::BruteForce` / `sequential::BarnesHut` over
 *     `Between(affected, affecting)` with `Acceleration<CHECKED>` (softening = 0) or
 *     `AccelerationSoftened<S, CHECKED>` (reference gravity/newtonian/acceleration*.rs):
 *         d = p2 - p1;  n = |d|^2;  CHECKED && n == 0 -> contributes nothing;
 *         otherwise a += d * mu2 / ((n + eps^2) * sqrt(n + eps^2))
 *     Empty affected -> nothing written; empty affecting -> zeros (the CPU paths' behaviour;
 *     the wgpu path panics on empty input, gpu/resources.rs:24).
 */
#ifndef PARTICULAR_CUDA_H
#define PARTICULAR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCUDA_ABI_VERSION 1

typedef enum pcuda_status {
    PCUDA_OK = 0,
    PCUDA_ERR_INVALID_ARGUMENT = -1,
    PCUDA_ERR_NO_DEVICE = -2,        /* no CUDA device / not compute capability 10.x */
    PCUDA_ERR_CUDA = -3,             /* a CUDA runtime call failed (see pcuda_last_error) */
    PCUDA_ERR_OUT_OF_MEMORY = -4,
    PCUDA_ERR_NCCL = -5,             /* NCCL missing or a collective failed */
    PCUDA_ERR_TREE_OVERFLOW = -6,    /* traversal stack / node capacity exceeded */
    PCUDA_ERR_NOT_INITIALISED = -7
} pcuda_status;

typedef struct pcuda_ctx pcuda_ctx;   /* opaque: device, stream, buffers, (optional) NCCL comm */
typedef struct pcuda_tree pcuda_tree; /* opaque: a built Barnes-Hut tree living on the device  */

typedef struct pcuda_config {
    int32_t device;        /* CUDA ordinal */
    uint32_t flags;        /* PCUDA_FLAG_* */
    uint32_t leaf_size;    /* max particles per Barnes-Hut leaf; 0 = default (16) */
    uint32_t expansion_order; /* Barnes-Hut node expansion: 0 or 1 = centre of mass (the reference's
                               * nodes, gravity/impls/mod.rs:103-135); 2 = + traceless quadrupole
                               * (an accuracy / speed knob beyond the reference, SURVEY.md 8f rank 4;
                               * f32 entry points only) */
} pcuda_config;

#define PCUDA_FLAG_NONE 0u
/* Do not record the per-phase CUDA events (pcuda_timings then reports only kernel_launches): saves
 * ~10 us per call, which matters at the reference's criterion sizes (N <= 65536). */
#define PCUDA_FLAG_NO_PHASE_TIMINGS 1u
/* Multi-GPU Barnes-Hut (pcuda_barneshut_f32x3_sharded*): how the tree is built.
 * PARTITIONED: every GPU sorts and builds only the tree of its own key range, the per-GPU trees
 * are exchanged and joined by a small top tree (SURVEY.md 8e "v3").  REPLICATED: every GPU builds
 * the whole tree (8e "v1").  Neither flag (and too few particles for LET, below): partitioned from
 * 4 GPUs on (measured on 8 B200s, N = 10M: 5.8 ms against 6.4 ms per step; N = 80M: 39.7 against
 * 44.6 ms; equal at 2 GPUs). */
#define PCUDA_FLAG_BH_PARTITIONED_BUILD 2u
#define PCUDA_FLAG_BH_REPLICATED_BUILD 4u
/* LET: locally essential trees.  The particles travel once, to the rank that owns their key range; every
 * rank builds the tree of its range and sends each other rank only the nodes and leaf particles that
 * rank's walk can open (the rest as stubs); cells that straddle a range boundary are joined by the same
 * top tree as in the partitioned build.  Nothing is replicated.  Default from 3 GPUs on when every rank
 * gets at least 65536 particles (N = 10M, ms per step, LET / partitioned / replicated: 2 GPUs 14.97 / 14.93 /
 * 14.86, 4 GPUs 8.18 / 8.34 / 8.93, 8 GPUs 5.01 / 5.31 / 6.04); the flag forces it. */
#define PCUDA_FLAG_BH_LET_BUILD 16u
/* `checked` with zero softening (Acceleration::checked, impls/mod.rs:160-161: a pair at zero distance
 * contributes nothing).  Small f32 brute-force problems (fewer than 2.5e8 pairs) and every f64 path
 * test r^2 == 0 exactly, as the reference does.  Large f32 brute-force problems instead add a floor
 *     t = 2 (max|mu| * 1e-38)^(2/3)      (max over the affecting set of the call; 9.3e-20 for mu = 1e9)
 * to every r^2, through the addend of the FMA chain that already adds softening^2 (no extra
 * instruction per pair): a coincident pair gives d * finite = 0 exactly, and r^2 + t == r^2 bit for bit
 * whenever r^2 >= 2^24 t.  That path therefore equals AccelerationSoftened::checked(sqrt(t)) — a
 * softening length of 3e-10 (max|mu| / 1e9)^(1/3) — and differs from Acceleration::checked() only for
 * pairs closer than 2^12 sqrt(t) = 1.2e-6 (max|mu| / 1e9)^(1/3), by a relative 1.5 t / r^2 of that pair's
 * term (above the 1e-5 parity bound only below 1.2e-7 (max|mu| / 1e9)^(1/3)).  The Barnes-Hut traversals
 * use the same floor with max|mu| replaced by n * max|mu| (a bound on every node mass).
 * EXACT_CHECKED: test r^2 == 0 exactly at every size in the f32 brute-force kernels (two more ALU
 * instructions per pair, about 5 % slower at N = 1M). */
#define PCUDA_FLAG_EXACT_CHECKED 8u

/* Per-phase device times of the LAST call on the context, in milliseconds (CUDA events on the
 * context stream).  Phases that did not run are 0.  Replaces nothing in the reference (it has no
 * metrics); used by bench.py. */
typedef struct pcuda_timings {
    float upload_ms;    /* host -> device copies                      */
    float comm_ms;      /* NCCL all-gather of sources (multi-GPU)     */
    float build_ms;     /* bbox + keys + sort + tree + centre of mass */
    float compute_ms;   /* brute-force kernel(s) or theta-traversal   */
    float download_ms;  /* device -> host copy of the result          */
    uint32_t kernel_launches; /* kernels of this library launched by the call */
    uint32_t reserved;
} pcuda_timings;

typedef struct pcuda_tree_info {
    uint64_t n_particles;
    uint64_t n_nodes;
    uint32_t n_levels;     /* root = level 0 */
    uint32_t leaf_size;
    uint32_t dim;
    uint32_t bits;         /* quantisation bits per axis: 21 (3-D) / 31 (2-D) */
    float origin[3];       /* min corner of the root cube (reference square_with) */
    float extent;          /* edge of the root cube */
    float inv;             /* 2^bits / extent */
    uint32_t reserved;
} pcuda_tree_info;

typedef enum pcuda_tree_array {
    PCUDA_TREE_KEYS = 0,        /* uint64[n]   sorted Morton keys                      */
    PCUDA_TREE_PERM = 1,        /* uint32[n]   sorted position -> original index       */
    PCUDA_TREE_NODE_BEGIN = 2,  /* uint32[m]   first sorted particle of the node       */
    PCUDA_TREE_NODE_COUNT = 3,  /* uint32[m]   particles in the node                   */
    PCUDA_TREE_NODE_LEVEL = 4,  /* uint32[m]                                           */
    PCUDA_TREE_NODE_FIRST_CHILD = 5, /* uint32[m] (0 for leaves)                       */
    PCUDA_TREE_NODE_NUM_CHILDREN = 6,/* uint32[m] (0 for leaves)                       */
    PCUDA_TREE_NODE_COM_MASS = 7     /* float[m][dim+1] {centre of mass, total mu}     */
} pcuda_tree_array;

/* ---- library / context ---------------------------------------------------------------------- */
int pcuda_abi_version(void);
const char *pcuda_status_string(int status);
int pcuda_device_count(int *count);

/* Replaces GpuResources::new + lazy WgpuResources::new (gpu/mod.rs:85-143, resources.rs:126-233):
 * create once, reuse across calls. */
int pcuda_create(const pcuda_config *config, pcuda_ctx **out);
void pcuda_destroy(pcuda_ctx *ctx);
/* Message of the last failed call on ctx (ctx == NULL: last failed pcuda_create on this thread). */
const char *pcuda_last_error(const pcuda_ctx *ctx);
int pcuda_get_timings(const pcuda_ctx *ctx, pcuda_timings *out);
/* cudaStream_t of the context (as void*), for interop with the *_dev entry points. */
void *pcuda_stream(pcuda_ctx *ctx);
/* Waits for the context stream; also folds the phase events of the last *_dev call into the
 * timings returned by pcuda_get_timings(). */
int pcuda_sync(pcuda_ctx *ctx);
/* Device properties the roofline needs: SM count, max SM clock (kHz). */
int pcuda_device_info(const pcuda_ctx *ctx, int *sm_count, int *sm_clock_khz, char *name,
                      size_t name_len);

/* Pinned host staging memory.  The reference packs particles straight into a mapped staging view
 * (T::write_affected(affected, view), gpu/mod.rs:187-195); the Rust wrapper packs into these
 * buffers so that uploads are true asynchronous DMA. */
int pcuda_host_alloc(pcuda_ctx *ctx, size_t bytes, void **out);
int pcuda_host_free(pcuda_ctx *ctx, void *p);

/* ---- brute force: replaces gpu::BruteForce::compute (gpu/mod.rs:179-208) ----------------------
 * HOST buffers in, HOST buffer out; blocking (the reference blocks too: pollster::block_on +
 * device.poll(Wait), gpu/mod.rs:200, resources.rs:340).  Upload, kernel, download. */
int pcuda_bruteforce_f32x3(pcuda_ctx *ctx, const float *affected_xyz, size_t n_affected,
                           const float *affecting_xyzm, size_t n_affecting, float softening,
                           int checked, float *out_xyz);
int pcuda_bruteforce_f32x2(pcuda_ctx *ctx, const float *affected_xy, size_t n_affected,
                           const float *affecting_xym, size_t n_affecting, float softening,
                           int checked, float *out_xy);
int pcuda_bruteforce_f64x3(pcuda_ctx *ctx, const double *affected_xyz, size_t n_affected,
                           const double *affecting_xyzm, size_t n_affecting, double softening,
                           int checked, double *out_xyz);
/* DVec2 (the pair term is wired for it at gravity/impls/glam.rs:231-235): {x,y,mu} records. */
int pcuda_bruteforce_f64x2(pcuda_ctx *ctx, const double *affected_xy, size_t n_affected,
                           const double *affecting_xym, size_t n_affecting, double softening,
                           int checked, double *out_xy);

/* DEVICE buffers in/out, enqueued on the context stream, returns without synchronising.
 * The device-resident stepping path and the multi-GPU driver use these. */
int pcuda_bruteforce_f32x3_dev(pcuda_ctx *ctx, const float *d_affected_xyz, size_t n_affected,
                               const float *d_affecting_xyzm, size_t n_affecting, float softening,
                               int checked, float *d_out_xyz);
int pcuda_bruteforce_f32x2_dev(pcuda_ctx *ctx, const float *d_affected_xy, size_t n_affected,
                               const float *d_affecting_xym, size_t n_affecting, float softening,
                               int checked, float *d_out_xy);
int pcuda_bruteforce_f64x3_dev(pcuda_ctx *ctx, const double *d_affected_xyz, size_t n_affected,
                               const double *d_affecting_xyzm, size_t n_affecting,
                               double softening, int checked, double *d_out_xyz);
int pcuda_bruteforce_f64x2_dev(pcuda_ctx *ctx, const double *d_affected_xy, size_t n_affected,
                               const double *d_affecting_xym, size_t n_affecting,
                               double softening, int checked, double *d_out_xy);

/* ---- Barnes-Hut: new on the GPU (the reference has sequential/parallel CPU versions only:
 * sequential.rs:439-543, parallel.rs:297-367).  One call = build the tree over `affecting`
 * (rebuilt every call like the reference, sequential.rs:539-541), then theta-traverse for every
 * affected particle.
 * `checked` is accepted for symmetry with the brute-force entry points and IGNORED: every Barnes-Hut
 * traversal is checked (a pair at zero distance contributes nothing; f32: through the r^2 floor
 * described at PCUDA_FLAG_EXACT_CHECKED, f64: exact test), because an unchecked tree walk has no
 * defined reference behaviour worth reproducing (NaN for every target that is also a source).
 * Deliberate deviation from sequential.rs:485-487: the reference SKIPS a node, internal or not, whose
 * centre of mass coincides with the target; here such a node is opened (theta^2 * 0 < w^2), which sums
 * its content exactly instead of dropping it.
 * Depth: cells are resolved down to extent / 2^21 (3-D) or extent / 2^31 (2-D), the resolution of the
 * Morton key; the reference subdivides until positions differ (tree/mod.rs:112-134).  Particles
 * closer than that share one leaf of unbounded size that is summed directly when opened: results
 * stay correct, but a cluster of k such particles costs O(k^2). */
int pcuda_barneshut_f32x3(pcuda_ctx *ctx, const float *affected_xyz, size_t n_affected,
                          const float *affecting_xyzm, size_t n_affecting, float theta,
                          float softening, int checked, float *out_xyz);
int pcuda_barneshut_f32x2(pcuda_ctx *ctx, const float *affected_xy, size_t n_affected,
                          const float *affecting_xym, size_t n_affecting, float theta,
                          float softening, int checked, float *out_xy);
int pcuda_barneshut_f32x3_dev(pcuda_ctx *ctx, const float *d_affected_xyz, size_t n_affected,
                              const float *d_affecting_xyzm, size_t n_affecting, float theta,
                              float softening, int checked, float *d_out_xyz);
int pcuda_barneshut_f32x2_dev(pcuda_ctx *ctx, const float *d_affected_xy, size_t n_affected,
                              const float *d_affecting_xym, size_t n_affecting, float theta,
                              float softening, int checked, float *d_out_xy);

/* Double precision (DVec3 / DVec2; the reference's BarnesHut is generic over the scalar,
 * sequential.rs:439-543, and its own tests run it in f64, gravity/newtonian/mod.rs:407-418).  The
 * tree STRUCTURE (keys, cells, opening decisions) is the f32 one over the particles rounded to
 * f32; sources, centres of mass, targets and the pair term are double precision, so theta = 0 is
 * the f64 brute-force sum.  Same argument meaning as the f32 entry points. */
int pcuda_barneshut_f64x3(pcuda_ctx *ctx, const double *affected_xyz, size_t n_affected,
                          const double *affecting_xyzm, size_t n_affecting, double theta,
                          double softening, int checked, double *out_xyz);
int pcuda_barneshut_f64x2(pcuda_ctx *ctx, const double *affected_xy, size_t n_affected,
                          const double *affecting_xym, size_t n_affecting, double theta,
                          double softening, int checked, double *out_xy);
int pcuda_barneshut_f64x3_dev(pcuda_ctx *ctx, const double *d_affected_xyz, size_t n_affected,
                              const double *d_affecting_xyzm, size_t n_affecting, double theta,
                              double softening, int checked, double *d_out_xyz);
int pcuda_barneshut_f64x2_dev(pcuda_ctx *ctx, const double *d_affected_xy, size_t n_affected,
                              const double *d_affecting_xym, size_t n_affecting, double theta,
                              double softening, int checked, double *d_out_xy);

/* Split phase (replaces RootedOrthtree::new, storage.rs:20-33, and
 * BarnesHut::compute(Between<&[P1], &RootedOrthtree>), sequential.rs:508-524): build once,
 * traverse many times, inspect the arrays for parity tests. HOST buffers. `dim` is 2 or 3. */
int pcuda_tree_build_f32(pcuda_ctx *ctx, uint32_t dim, const float *affecting, size_t n_affecting,
                         pcuda_tree **out);
int pcuda_tree_info_get(const pcuda_tree *tree, pcuda_tree_info *out);
int pcuda_tree_read(pcuda_ctx *ctx, const pcuda_tree *tree, int which /* pcuda_tree_array */,
                    void *dst, size_t dst_bytes);
int pcuda_tree_traverse_f32(pcuda_ctx *ctx, const pcuda_tree *tree, const float *affected,
                            size_t n_affected, float theta, float softening, int checked,
                            float *out);
/* Instrumentation of the LAST traversal: counters[0] = accepted node (centre-of-mass)
 * interactions and [1] = direct particle interactions, both summed over all targets;
 * [2] = node tests and [3] = interaction-list entries (nodes + particles), both summed over the
 * <= 32-target groups that walk the tree together; [4] = number of groups. */
int pcuda_tree_last_counters(pcuda_ctx *ctx, uint64_t counters[5]);
void pcuda_tree_destroy(pcuda_ctx *ctx, pcuda_tree *tree);

/* Morton keys and the stable sort permutation alone (the first half of the build; SURVEY.md 8b/8c:
 * the reference has no Morton code, tree/mod.rs:117-126 is a recursive bucket partition, so the
 * specification is ours — root cube per BoundingBox::square_with, tree/partition.rs:136-153,
 * 21 bits per axis in 3-D / 31 in 2-D, axis 0 in the lowest bit — and parity is bit-exact against
 * oracle/oracle_octree.inc).  HOST buffers: particles = n records {x,y[,z],mu};
 * keys_out[i] = i-th smallest key, perm_out[i] = input index of the particle holding it (equal keys
 * keep input order).  frame_out (may be NULL) receives origin / extent / inv of the root cube. */
int pcuda_morton_f32x3(pcuda_ctx *ctx, const float *particles_xyzm, size_t n, uint64_t *keys_out,
                       uint32_t *perm_out, pcuda_tree_info *frame_out);
int pcuda_morton_f32x2(pcuda_ctx *ctx, const float *particles_xym, size_t n, uint64_t *keys_out,
                       uint32_t *perm_out, pcuda_tree_info *frame_out);

/* ---- multi-GPU (one process per GPU; new — the reference is single-device) --------------------
 * Targets are sharded by the caller; sources are replicated with an all-gather over NVLink each
 * step.  NCCL is dlopen()ed on first use.  id is the 128-byte ncclUniqueId made by rank 0 and
 * distributed by the host side (torch.distributed / MPI / a socket). */
#define PCUDA_UNIQUE_ID_BYTES 128
int pcuda_comm_unique_id(pcuda_ctx *ctx, uint8_t id[PCUDA_UNIQUE_ID_BYTES]);
int pcuda_comm_init(pcuda_ctx *ctx, const uint8_t id[PCUDA_UNIQUE_ID_BYTES], int world_size,
                    int rank);
int pcuda_comm_destroy(pcuda_ctx *ctx);
/* In-process communicator for TESTS: binds `world_size` contexts of this process (rank = index; one
 * host thread drives each; same device, or devices with peer access) so that the sharded entry points
 * below run all their ranks on a box with a single GPU: the collectives become device-to-device
 * copies ordered by events behind the same internal calls as NCCL.  Call once, from one thread,
 * before the rank threads start; every collective is a rendezvous of all ranks. */
int pcuda_comm_init_local(pcuda_ctx *const *ctxs, int world_size);
/* All-gather equally sized shards of `bytes_per_rank` bytes (device pointers, context stream). */
int pcuda_comm_allgather_dev(pcuda_ctx *ctx, const void *d_send, void *d_recv,
                             size_t bytes_per_rank);

/* One multi-GPU brute-force step, device-resident (new).  Every rank passes the
 * {x,y,z,mu} records of the particles it owns (n_local <= shard_capacity; shard_capacity must be
 * the same on all ranks).  The records are placed in this rank's slot of d_gathered_xyzm
 * (world_size * shard_capacity records; unused slots are padded with zero-mass records that
 * contribute exactly 0), all-gathered in place, and the local particles are evaluated against
 * all of them: d_out_xyz[i] = acceleration of local particle i (n_local x 3).  Without a
 * communicator (pcuda_comm_init not called) this is the single-GPU all-pairs evaluation.
 * Enqueued on the context stream; does not synchronise; pcuda_get_timings() after pcuda_sync()
 * does not include this call's phases (use events on pcuda_stream()). */
int pcuda_bruteforce_f32x3_sharded_dev(pcuda_ctx *ctx, const float *d_local_xyzm, size_t n_local,
                                       size_t shard_capacity, float softening, int checked,
                                       float *d_gathered_xyzm, float *d_out_xyz);

/* The same step with HOST buffers (upload of the local records, step, download of the local
 * accelerations); blocking; fills pcuda_get_timings() including comm_ms. */
int pcuda_bruteforce_f32x3_sharded(pcuda_ctx *ctx, const float *local_xyzm, size_t n_local,
                                   size_t shard_capacity, float softening, int checked,
                                   float *out_xyz);

/* Multi-GPU `Between(affected, affecting)` step for the massive -> massless split (BASELINE
 * configs[2]; the storage mapping of Reordered, storage.rs:153-163, 207-229):
 * the AFFECTED positions are sharded by the caller (this rank passes only its n_affected targets —
 * no collective ever touches them), the AFFECTING records are sharded too (n_local_src <=
 * src_capacity per rank, same capacity on all ranks) and all-gathered in place into
 * d_gathered_src_xyzm (world_size * src_capacity records, zero-mass padding contributes exactly 0).
 * out[i] = acceleration of this rank's i-th affected particle.  Every rank must call, also with
 * n_affected == 0.  _dev: device pointers, enqueued on the context stream, no synchronisation.
 * Host version: blocking; large target shards take the chunked upload / evaluate / download
 * pipeline of pcuda_bruteforce_f32x3. */
int pcuda_bruteforce_f32x3_between_sharded_dev(pcuda_ctx *ctx, const float *d_affected_xyz,
                                               size_t n_affected, const float *d_local_src_xyzm,
                                               size_t n_local_src, size_t src_capacity,
                                               float softening, int checked,
                                               float *d_gathered_src_xyzm, float *d_out_xyz);
int pcuda_bruteforce_f32x3_between_sharded(pcuda_ctx *ctx, const float *affected_xyz,
                                           size_t n_affected, const float *local_src_xyzm,
                                           size_t n_local_src, size_t src_capacity, float softening,
                                           int checked, float *out_xyz);

/* One multi-GPU Barnes-Hut step ("replicated build"): rank r owns the contiguous
 * block [r * cap, r * cap + n_local) of the n_total particles, cap = ceil(n_total / world_size)
 * (the call fails if n_local does not match).  The records are all-gathered in place into
 * d_gathered_xyzm (world_size * cap records) and every GPU builds the identical tree over all
 * n_total particles.  The traversal is sharded by key range (rank r walks the tree for the r-th
 * block of the Morton-sorted particles: compact target groups, no target sort), the per-range
 * accelerations are all-gathered (12 B per particle) and every rank keeps the rows of the
 * particles it owns: out[i] = acceleration of local particle i.
 * _dev: device pointers, enqueued on the context stream, no synchronisation. */
int pcuda_barneshut_f32x3_sharded_dev(pcuda_ctx *ctx, const float *d_local_xyzm, size_t n_local,
                                      size_t n_total, float theta, float softening, int checked,
                                      float *d_gathered_xyzm, float *d_out_xyz);
int pcuda_barneshut_f32x3_sharded(pcuda_ctx *ctx, const float *local_xyzm, size_t n_local,
                                  size_t n_total, float theta, float softening, int checked,
                                  float *out_xyz);

/* Partitioned build (PCUDA_FLAG_BH_PARTITIONED_BUILD): the sharded entry points above cut the key
 * space into world_size ranges of about equal population (quantiles of a regular sample of the
 * keys, the same on every rank), rank r selects, sorts and builds the tree of the r-th range only —
 * same root cube, same level and leaf rules — the node records and sort permutations are
 * all-gathered into equal slots, the cells that straddle a range boundary (at most the first and
 * the last node of every level of every part) are merged into a small top tree (moments added in
 * double precision), and every rank walks the joined tree for the targets of its own key range:
 * the same cells and centres of mass as the single tree, up to the order of the f64 additions.
 * The two entry points below run the same partition / per-part build / forest walk on ONE GPU,
 * part after part, for all particles ("virtual ranks"): the test vehicle of the multi-GPU path.
 * xyzm: n {x,y,z,mu} records; out: n accelerations in input order; 1 <= parts <= 16;
 * parts == 1 gives the ordinary tree. */
int pcuda_barneshut_f32x3_partitioned(pcuda_ctx *ctx, const float *xyzm, size_t n, int parts,
                                      float theta, float softening, int checked, float *out_xyz);
int pcuda_barneshut_f32x3_partitioned_dev(pcuda_ctx *ctx, const float *d_xyzm, size_t n, int parts,
                                          float theta, float softening, int checked,
                                          float *d_out_xyz);

/* ---- device-resident stepping (new) --------------------------------------
 * Every caller of the reference integrates the accelerations right after computing them
 * (examples/simple/src/main.rs:45-59, examples/particle-toy/src/physics.rs:128-138 + 166-176, the
 * reference's own circular_orbit! test gravity/newtonian/mod.rs:318-331), and its wgpu operator
 * uploads and reads back on every call (gpu/resources.rs:37-39, 318-349).  A pcuda_sim keeps
 * particles {position, mu}, velocities and accelerations in device memory; one step is
 *     a = algorithm(Between(all particles, affecting particles))
 *     velocity += a * dt;  position += velocity * dt          (semi-implicit Euler, unfused)
 * and nothing crosses PCIe until pcuda_sim_read().  Brute-force steps replay from a CUDA graph. */
typedef struct pcuda_sim pcuda_sim;
typedef enum pcuda_algorithm { PCUDA_BRUTE_FORCE = 0, PCUDA_BARNES_HUT = 1 } pcuda_algorithm;
typedef enum pcuda_scalar { PCUDA_F32 = 0, PCUDA_F64 = 1 } pcuda_scalar;
/* affecting = the particles with mu != 0, in input order; affected = all particles in input order:
 * the `Reordered` storage (storage.rs:153-163, 219-229; ring-formation/src/nbody.rs:25-28). */
#define PCUDA_SIM_AFFECTING_MASSIVE_ONLY 1u
#define PCUDA_SIM_NO_GRAPH 2u /* launch every kernel individually (debugging / profiling) */

typedef struct pcuda_sim_config {
    uint32_t dim;       /* 2 or 3 */
    uint32_t scalar;    /* pcuda_scalar */
    uint32_t algorithm; /* pcuda_algorithm */
    uint32_t flags;     /* PCUDA_SIM_* */
    double theta;       /* Barnes-Hut opening parameter */
    double softening;   /* 0 = Acceleration, > 0 = AccelerationSoftened */
    double dt;
    int32_t checked;    /* brute force: a zero-distance pair contributes nothing */
    uint32_t reserved;
} pcuda_sim_config;

typedef struct pcuda_sim_info_t {
    uint64_t n_particles;
    uint64_t n_affecting;
    uint64_t steps_done;
    void *d_particles;     /* device: n x (dim+1) scalars {position, mu} — e.g. for rendering */
    void *d_velocities;    /* device: n x dim */
    void *d_accelerations; /* device: n x dim, accelerations of the last step */
    uint32_t graph_active; /* brute-force steps are being replayed from a CUDA graph */
    uint32_t launches_per_step;
} pcuda_sim_info_t;

/* HOST buffers: particles n x (dim+1), velocities n x dim (NULL = at rest); copied, not retained. */
int pcuda_sim_create(pcuda_ctx *ctx, const pcuda_sim_config *config, const void *particles,
                     const void *velocities, size_t n, pcuda_sim **out);
/* Changes theta / softening / dt / checked / algorithm of a live simulation. */
int pcuda_sim_configure(pcuda_ctx *ctx, pcuda_sim *sim, const pcuda_sim_config *config);
/* Enqueues n_steps steps on the context stream; does not wait for them. */
int pcuda_sim_step(pcuda_ctx *ctx, pcuda_sim *sim, uint32_t n_steps);
/* Blocking read-back into HOST buffers; any of the three may be NULL. */
int pcuda_sim_read(pcuda_ctx *ctx, pcuda_sim *sim, void *particles, void *velocities,
                   void *accelerations);
int pcuda_sim_info(const pcuda_sim *sim, pcuda_sim_info_t *out);
void pcuda_sim_destroy(pcuda_ctx *ctx, pcuda_sim *sim);

/* ---- user-defined interactions (new) -------------------------------------
 * The reference's GPU operator is generic over the interaction: an InteractionShader supplies WGSL
 * source defining the types Affected / Affecting / Interaction and `fn compute(p1, p2, out)`, their
 * byte sizes and optional push constants (gpu/mod.rs:40-82), which the operator pastes into a
 * brute-force template and compiles at run time (gpu/bruteforce.wgsl, gpu/resources.rs:126-233).
 * CUDA counterpart: `source` is CUDA C++ that defines
 *     struct Affected {..}; struct Affecting {..}; struct Interaction {..}; struct Push {..};
 *     __device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out);
 * (`push` is a __constant__ Push visible to compute; Interaction() is the start value, as
 * `var out = Interaction()` in the WGSL template; struct sizes must be multiples of 4 bytes; floating
 * point contraction is off, so a * b + c rounds twice like the Rust CPU path unless the source calls
 * fmaf itself).  It is compiled with NVRTC for sm_100a and evaluated as
 *     interactions[i] = fold over all affecting j, in order, of compute(affected[i], affecting[j], out)
 * Gravity itself does not go through here (Acceleration / AccelerationSoftened have tuned kernels). */
typedef struct pcuda_interaction pcuda_interaction;
/* Compile only (no device needed): status + compiler log; the analogue of shader validation. */
int pcuda_interaction_check(const char *source, char *log, size_t log_len);
int pcuda_interaction_create(pcuda_ctx *ctx, const char *source, pcuda_interaction **out);
/* sizes[0..3] = sizeof(Affected), sizeof(Affecting), sizeof(Interaction), sizeof(Push) on the device
 * (AFFECTED_SIZE / AFFECTING_SIZE / INTERACTION_SIZE of the reference trait). */
int pcuda_interaction_sizes(const pcuda_interaction *interaction, uint32_t sizes[4]);
/* HOST buffers laid out as arrays of the source's structs; blocking. */
int pcuda_interaction_brute_force(pcuda_ctx *ctx, pcuda_interaction *interaction, const void *affected,
                                  size_t n_affected, const void *affecting, size_t n_affecting,
                                  const void *push, size_t push_bytes, void *out);
void pcuda_interaction_destroy(pcuda_ctx *ctx, pcuda_interaction *interaction);

#ifdef __cplusplus
}
#endif
#endif /* PARTICULAR_CUDA_H */
