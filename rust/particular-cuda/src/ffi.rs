//! `extern "C"` declarations, 1:1 with `include/particular_cuda.h` (ABI version 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct pcuda_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pcuda_tree {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pcuda_sim {
    _private: [u8; 0],
}
#[repr(C)]
pub struct pcuda_interaction {
    _private: [u8; 0],
}

pub const PCUDA_OK: c_int = 0;
pub const PCUDA_ERR_NO_DEVICE: c_int = -2;
pub const PCUDA_UNIQUE_ID_BYTES: usize = 128;
pub const PCUDA_FLAG_NO_PHASE_TIMINGS: u32 = 1;
pub const PCUDA_FLAG_BH_PARTITIONED_BUILD: u32 = 2;
pub const PCUDA_FLAG_BH_REPLICATED_BUILD: u32 = 4;
pub const PCUDA_FLAG_EXACT_CHECKED: u32 = 8;
pub const PCUDA_FLAG_BH_LET_BUILD: u32 = 16;

pub const PCUDA_BRUTE_FORCE: u32 = 0;
pub const PCUDA_BARNES_HUT: u32 = 1;
pub const PCUDA_F32: u32 = 0;
pub const PCUDA_F64: u32 = 1;
pub const PCUDA_SIM_AFFECTING_MASSIVE_ONLY: u32 = 1;
pub const PCUDA_SIM_NO_GRAPH: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct pcuda_config {
    pub device: i32,
    pub flags: u32,
    pub leaf_size: u32,
    pub expansion_order: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct pcuda_timings {
    pub upload_ms: f32,
    pub comm_ms: f32,
    pub build_ms: f32,
    pub compute_ms: f32,
    pub download_ms: f32,
    pub kernel_launches: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct pcuda_tree_info {
    pub n_particles: u64,
    pub n_nodes: u64,
    pub n_levels: u32,
    pub leaf_size: u32,
    pub dim: u32,
    pub bits: u32,
    pub origin: [f32; 3],
    pub extent: f32,
    pub inv: f32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct pcuda_sim_config {
    pub dim: u32,
    pub scalar: u32,
    pub algorithm: u32,
    pub flags: u32,
    pub theta: f64,
    pub softening: f64,
    pub dt: f64,
    pub checked: i32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct pcuda_sim_info_t {
    pub n_particles: u64,
    pub n_affecting: u64,
    pub steps_done: u64,
    pub d_particles: *mut c_void,
    pub d_velocities: *mut c_void,
    pub d_accelerations: *mut c_void,
    pub graph_active: u32,
    pub launches_per_step: u32,
}

extern "C" {
    pub fn pcuda_abi_version() -> c_int;
    pub fn pcuda_status_string(status: c_int) -> *const c_char;
    pub fn pcuda_device_count(count: *mut c_int) -> c_int;
    pub fn pcuda_create(config: *const pcuda_config, out: *mut *mut pcuda_ctx) -> c_int;
    pub fn pcuda_destroy(ctx: *mut pcuda_ctx);
    pub fn pcuda_last_error(ctx: *const pcuda_ctx) -> *const c_char;
    pub fn pcuda_get_timings(ctx: *const pcuda_ctx, out: *mut pcuda_timings) -> c_int;
    pub fn pcuda_stream(ctx: *mut pcuda_ctx) -> *mut c_void;
    pub fn pcuda_sync(ctx: *mut pcuda_ctx) -> c_int;
    pub fn pcuda_device_info(ctx: *const pcuda_ctx, sm_count: *mut c_int, sm_clock_khz: *mut c_int,
                             name: *mut c_char, name_len: usize) -> c_int;
    pub fn pcuda_host_alloc(ctx: *mut pcuda_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn pcuda_host_free(ctx: *mut pcuda_ctx, p: *mut c_void) -> c_int;

    pub fn pcuda_bruteforce_f32x3(ctx: *mut pcuda_ctx, affected_xyz: *const f32, n_affected: usize,
                                  affecting_xyzm: *const f32, n_affecting: usize, softening: f32,
                                  checked: c_int, out_xyz: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f32x2(ctx: *mut pcuda_ctx, affected_xy: *const f32, n_affected: usize,
                                  affecting_xym: *const f32, n_affecting: usize, softening: f32,
                                  checked: c_int, out_xy: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f64x3(ctx: *mut pcuda_ctx, affected_xyz: *const f64, n_affected: usize,
                                  affecting_xyzm: *const f64, n_affecting: usize, softening: f64,
                                  checked: c_int, out_xyz: *mut f64) -> c_int;
    pub fn pcuda_bruteforce_f64x2(ctx: *mut pcuda_ctx, affected_xy: *const f64, n_affected: usize,
                                  affecting_xym: *const f64, n_affecting: usize, softening: f64,
                                  checked: c_int, out_xy: *mut f64) -> c_int;
    pub fn pcuda_bruteforce_f64x2_dev(ctx: *mut pcuda_ctx, d_affected: *const f64, n_affected: usize,
                                      d_affecting: *const f64, n_affecting: usize, softening: f64,
                                      checked: c_int, d_out: *mut f64) -> c_int;
    pub fn pcuda_bruteforce_f32x3_dev(ctx: *mut pcuda_ctx, d_affected: *const f32, n_affected: usize,
                                      d_affecting: *const f32, n_affecting: usize, softening: f32,
                                      checked: c_int, d_out: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f32x2_dev(ctx: *mut pcuda_ctx, d_affected: *const f32, n_affected: usize,
                                      d_affecting: *const f32, n_affecting: usize, softening: f32,
                                      checked: c_int, d_out: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f64x3_dev(ctx: *mut pcuda_ctx, d_affected: *const f64, n_affected: usize,
                                      d_affecting: *const f64, n_affecting: usize, softening: f64,
                                      checked: c_int, d_out: *mut f64) -> c_int;

    pub fn pcuda_barneshut_f32x3(ctx: *mut pcuda_ctx, affected_xyz: *const f32, n_affected: usize,
                                 affecting_xyzm: *const f32, n_affecting: usize, theta: f32,
                                 softening: f32, checked: c_int, out_xyz: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x2(ctx: *mut pcuda_ctx, affected_xy: *const f32, n_affected: usize,
                                 affecting_xym: *const f32, n_affecting: usize, theta: f32,
                                 softening: f32, checked: c_int, out_xy: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x3_dev(ctx: *mut pcuda_ctx, d_affected: *const f32, n_affected: usize,
                                     d_affecting: *const f32, n_affecting: usize, theta: f32,
                                     softening: f32, checked: c_int, d_out: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x2_dev(ctx: *mut pcuda_ctx, d_affected: *const f32, n_affected: usize,
                                     d_affecting: *const f32, n_affecting: usize, theta: f32,
                                     softening: f32, checked: c_int, d_out: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f64x3(ctx: *mut pcuda_ctx, affected_xyz: *const f64, n_affected: usize,
                                 affecting_xyzm: *const f64, n_affecting: usize, theta: f64,
                                 softening: f64, checked: c_int, out_xyz: *mut f64) -> c_int;
    pub fn pcuda_barneshut_f64x2(ctx: *mut pcuda_ctx, affected_xy: *const f64, n_affected: usize,
                                 affecting_xym: *const f64, n_affecting: usize, theta: f64,
                                 softening: f64, checked: c_int, out_xy: *mut f64) -> c_int;
    pub fn pcuda_barneshut_f64x3_dev(ctx: *mut pcuda_ctx, d_affected: *const f64, n_affected: usize,
                                     d_affecting: *const f64, n_affecting: usize, theta: f64,
                                     softening: f64, checked: c_int, d_out: *mut f64) -> c_int;
    pub fn pcuda_barneshut_f64x2_dev(ctx: *mut pcuda_ctx, d_affected: *const f64, n_affected: usize,
                                     d_affecting: *const f64, n_affecting: usize, theta: f64,
                                     softening: f64, checked: c_int, d_out: *mut f64) -> c_int;

    pub fn pcuda_tree_build_f32(ctx: *mut pcuda_ctx, dim: u32, affecting: *const f32, n: usize,
                                out: *mut *mut pcuda_tree) -> c_int;
    pub fn pcuda_tree_info_get(tree: *const pcuda_tree, out: *mut pcuda_tree_info) -> c_int;
    pub fn pcuda_tree_read(ctx: *mut pcuda_ctx, tree: *const pcuda_tree, which: c_int, dst: *mut c_void,
                           dst_bytes: usize) -> c_int;
    pub fn pcuda_tree_traverse_f32(ctx: *mut pcuda_ctx, tree: *const pcuda_tree, affected: *const f32,
                                   n_affected: usize, theta: f32, softening: f32, checked: c_int,
                                   out: *mut f32) -> c_int;
    pub fn pcuda_tree_last_counters(ctx: *mut pcuda_ctx, counters: *mut u64) -> c_int;
    pub fn pcuda_tree_destroy(ctx: *mut pcuda_ctx, tree: *mut pcuda_tree);

    pub fn pcuda_comm_unique_id(ctx: *mut pcuda_ctx, id: *mut u8) -> c_int;
    pub fn pcuda_comm_init(ctx: *mut pcuda_ctx, id: *const u8, world_size: c_int, rank: c_int) -> c_int;
    pub fn pcuda_comm_destroy(ctx: *mut pcuda_ctx) -> c_int;
    pub fn pcuda_comm_init_local(ctxs: *const *mut pcuda_ctx, world_size: c_int) -> c_int;
    pub fn pcuda_comm_allgather_dev(ctx: *mut pcuda_ctx, d_send: *const c_void, d_recv: *mut c_void,
                                    bytes_per_rank: usize) -> c_int;
    pub fn pcuda_bruteforce_f32x3_sharded_dev(ctx: *mut pcuda_ctx, d_local_xyzm: *const f32, n_local: usize,
                                              shard_capacity: usize, softening: f32, checked: c_int,
                                              d_gathered_xyzm: *mut f32, d_out_xyz: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f32x3_sharded(ctx: *mut pcuda_ctx, local_xyzm: *const f32, n_local: usize,
                                          shard_capacity: usize, softening: f32, checked: c_int,
                                          out_xyz: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f32x3_between_sharded_dev(ctx: *mut pcuda_ctx, d_affected_xyz: *const f32,
                                                      n_affected: usize, d_local_src_xyzm: *const f32,
                                                      n_local_src: usize, src_capacity: usize, softening: f32,
                                                      checked: c_int, d_gathered_src_xyzm: *mut f32,
                                                      d_out_xyz: *mut f32) -> c_int;
    pub fn pcuda_bruteforce_f32x3_between_sharded(ctx: *mut pcuda_ctx, affected_xyz: *const f32,
                                                  n_affected: usize, local_src_xyzm: *const f32,
                                                  n_local_src: usize, src_capacity: usize, softening: f32,
                                                  checked: c_int, out_xyz: *mut f32) -> c_int;
    pub fn pcuda_morton_f32x3(ctx: *mut pcuda_ctx, particles_xyzm: *const f32, n: usize, keys_out: *mut u64,
                              perm_out: *mut u32, frame_out: *mut pcuda_tree_info) -> c_int;
    pub fn pcuda_morton_f32x2(ctx: *mut pcuda_ctx, particles_xym: *const f32, n: usize, keys_out: *mut u64,
                              perm_out: *mut u32, frame_out: *mut pcuda_tree_info) -> c_int;
    pub fn pcuda_barneshut_f32x3_sharded_dev(ctx: *mut pcuda_ctx, d_local_xyzm: *const f32, n_local: usize,
                                             n_total: usize, theta: f32, softening: f32, checked: c_int,
                                             d_gathered_xyzm: *mut f32, d_out_xyz: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x3_sharded(ctx: *mut pcuda_ctx, local_xyzm: *const f32, n_local: usize,
                                         n_total: usize, theta: f32, softening: f32, checked: c_int,
                                         out_xyz: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x3_partitioned(ctx: *mut pcuda_ctx, xyzm: *const f32, n: usize, parts: c_int,
                                             theta: f32, softening: f32, checked: c_int,
                                             out_xyz: *mut f32) -> c_int;
    pub fn pcuda_barneshut_f32x3_partitioned_dev(ctx: *mut pcuda_ctx, d_xyzm: *const f32, n: usize, parts: c_int,
                                                 theta: f32, softening: f32, checked: c_int,
                                                 d_out_xyz: *mut f32) -> c_int;

    pub fn pcuda_sim_create(ctx: *mut pcuda_ctx, config: *const pcuda_sim_config, particles: *const c_void,
                            velocities: *const c_void, n: usize, out: *mut *mut pcuda_sim) -> c_int;
    pub fn pcuda_sim_configure(ctx: *mut pcuda_ctx, sim: *mut pcuda_sim, config: *const pcuda_sim_config) -> c_int;
    pub fn pcuda_sim_step(ctx: *mut pcuda_ctx, sim: *mut pcuda_sim, n_steps: u32) -> c_int;
    pub fn pcuda_sim_read(ctx: *mut pcuda_ctx, sim: *mut pcuda_sim, particles: *mut c_void,
                          velocities: *mut c_void, accelerations: *mut c_void) -> c_int;
    pub fn pcuda_sim_info(sim: *const pcuda_sim, out: *mut pcuda_sim_info_t) -> c_int;
    pub fn pcuda_sim_destroy(ctx: *mut pcuda_ctx, sim: *mut pcuda_sim);

    pub fn pcuda_interaction_check(source: *const c_char, log: *mut c_char, log_len: usize) -> c_int;
    pub fn pcuda_interaction_create(ctx: *mut pcuda_ctx, source: *const c_char, out: *mut *mut pcuda_interaction) -> c_int;
    pub fn pcuda_interaction_sizes(interaction: *const pcuda_interaction, sizes: *mut u32) -> c_int;
    pub fn pcuda_interaction_brute_force(ctx: *mut pcuda_ctx, interaction: *mut pcuda_interaction,
                                         affected: *const c_void, n_affected: usize, affecting: *const c_void,
                                         n_affecting: usize, push: *const c_void, push_bytes: usize,
                                         out: *mut c_void) -> c_int;
    pub fn pcuda_interaction_destroy(ctx: *mut pcuda_ctx, interaction: *mut pcuda_interaction);
}
