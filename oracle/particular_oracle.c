/*
 * particular_oracle.c — CPU restatement of particular's brute-force and Barnes-Hut arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under particular_b200/ or include/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, and only as the checker / reported baseline.  The product path is the CUDA library and
 * fails loudly when it is missing.
 *
 * What it restates (paths relative to /root/reference/particular/src):
 *   pair term          gravity/impls/mod.rs:151-166
 *   left fold          sequential.rs:181-194, 101-106
 *   root cube          tree/partition.rs:84-92, 109-153
 *   recursive build    tree/mod.rs:91-138, storage.rs:20-33
 *   centre of mass     gravity/impls/mod.rs:120-134
 *   theta traversal    sequential.rs:466-505, gravity/newtonian/acceleration.rs:111-127
 *
 * Pinning: the reference is Rust and there is no cargo/rustc in this image, so the reference
 * itself cannot be executed here (no oracle/_ref).  The oracle is pinned instead against every
 * known-answer the reference's own tests hold for this path: the six-particle
 * `acceleration_error!` fixture with its closed form (gravity/newtonian/mod.rs:228-277), the
 * `circular_orbit!` drift bound (:281-347), the doctest force identities (lib.rs:247-261) and
 * "BarnesHut(theta = 0) == brute force" (:409-413) — see tests/test_oracle_golden.py.
 * The Morton-key / sort / linear-octree functions at the bottom restate OUR OWN specification
 * (DESIGN.md "Tree specification"): the reference contains no Morton code, so that part is
 * "parity unpinned" by the reference and pinned only CPU-vs-GPU.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define S float
#define D 3
#define SQRT sqrtf
#define FN(name) oracle_##name##_f32x3
#include "oracle_impl.inc"
#undef S
#undef D
#undef SQRT
#undef FN

#define S float
#define D 2
#define SQRT sqrtf
#define FN(name) oracle_##name##_f32x2
#include "oracle_impl.inc"
#undef S
#undef D
#undef SQRT
#undef FN

#define S double
#define D 3
#define SQRT sqrt
#define FN(name) oracle_##name##_f64x3
#include "oracle_impl.inc"
#undef S
#undef D
#undef SQRT
#undef FN

#define S double
#define D 2
#define SQRT sqrt
#define FN(name) oracle_##name##_f64x2
#include "oracle_impl.inc"
#undef S
#undef D
#undef SQRT
#undef FN

#define D 3
#define BITS 21
#define FN(name) oracle_##name##_f32x3
#include "oracle_octree.inc"
#undef D
#undef BITS
#undef FN

#define D 2
#define BITS 31
#define FN(name) oracle_##name##_f32x2
#include "oracle_octree.inc"
#undef D
#undef BITS
#undef FN

int oracle_abi_version(void) { return 1; }
