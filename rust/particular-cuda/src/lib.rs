//! # particular-cuda
//!
//! B200 (sm_100a) compute backend for [`particular`]: `cuda::BruteForce` and `cuda::BarnesHut` implement
//! the crate's operator trait `Interaction<Between<&[P1], &[P2]>>` the same way the wgpu operator
//! `gpu::BruteForce` does (particular/src/gpu/mod.rs:179-208), so every storage the crate knows
//! (`&[P]`, `&Ordered<P>`, `&Reordered<P, F>`, `Between<..>` — blanket impls at storage.rs:207-241)
//! and every `#[derive(Position, Mass)]` user type work unchanged:
//!
//! ```ignore
//! let mut ctx = particular_cuda::CudaContext::new(0);
//! // before: bodies.brute_force_simd::<8>(Acceleration::checked())
//! let accelerations = bodies.as_slice().cuda_brute_force(&mut ctx, Acceleration::checked());
//! let accelerations = bodies.as_slice().cuda_barnes_hut(&mut ctx, 0.5, Acceleration::checked());
//! ```
//!
//! All arithmetic happens in `libparticular_cuda.so` (hand-written CUDA, built by `build.rs` with
//! `nvcc -gencode arch=compute_100a,code=sm_100a`).  There is no CPU fallback: `CudaContext::new`
//! panics without a compute-capability-10 device, as the reference's wgpu path panics without an
//! adapter.
//!
//! **This crate has not been compiled in the repository's build environment (no Rust toolchain
//! there).**  The identical interface is built and tested in C++ (`include/particular_cuda.hpp`) and
//! Python (`particular_b200/interface.py`) over the same C ABI.
pub mod ffi;

#[cfg(feature = "compat-0-7")]
pub mod compat;

use std::ffi::CStr;
use std::marker::PhantomData;

use particular::gravity::newtonian::{Acceleration, AccelerationSoftened};
use particular::gravity::{Mass, Position};
use particular::storage::{Ordered, Reordered};
use particular::{Between, Interaction};

fn last_error(ctx: *const ffi::pcuda_ctx) -> String {
    // SAFETY: the library returns a NUL-terminated string owned by the context / thread.
    unsafe { CStr::from_ptr(ffi::pcuda_last_error(ctx)) }.to_string_lossy().into_owned()
}

#[track_caller]
fn check(status: i32, ctx: *const ffi::pcuda_ctx) {
    // The reference unwrap()s / expect()s on the wgpu path (gpu/resources.rs:38-39, 341-345).
    assert_eq!(status, ffi::PCUDA_OK, "particular-cuda: {}", last_error(ctx));
}

/// One device, one stream, grow-only device buffers and (optionally) an NCCL communicator: the
/// analogue of `GpuResources` + `wgpu::Device` + `wgpu::Queue` (gpu/mod.rs:85-159).  Create once and
/// reuse for every step ("should not be recreated for every iteration", gpu/mod.rs:150-151).
pub struct CudaContext {
    raw: *mut ffi::pcuda_ctx,
}

// One call in flight per context (`&mut CudaContext`), movable between threads, not shareable.
unsafe impl Send for CudaContext {}

impl CudaContext {
    /// Panics when `device` is not a usable sm_100 GPU (no CPU fallback).
    pub fn new(device: i32) -> Self {
        Self::with_leaf_size(device, 0)
    }

    /// `leaf_size`: largest Barnes-Hut leaf (0 = default 16).  Per-phase device timings are off
    /// (`PCUDA_FLAG_NO_PHASE_TIMINGS`): the product path records no events.
    pub fn with_leaf_size(device: i32, leaf_size: u32) -> Self {
        Self::with_flags(device, leaf_size, 0)
    }

    /// `flags`: `ffi::PCUDA_FLAG_BH_PARTITIONED_BUILD` / `ffi::PCUDA_FLAG_BH_REPLICATED_BUILD` force how
    /// multi-GPU Barnes-Hut builds its tree (default: partitioned by key range from 4 GPUs on);
    /// `ffi::PCUDA_FLAG_EXACT_CHECKED` makes the f32 brute-force kernels test `r^2 == 0` exactly at every size.
    pub fn with_flags(device: i32, leaf_size: u32, flags: u32) -> Self {
        let cfg = ffi::pcuda_config {
            device,
            flags: ffi::PCUDA_FLAG_NO_PHASE_TIMINGS | flags,
            leaf_size,
            expansion_order: 1,
        };
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::pcuda_create(&cfg, &mut raw) }, std::ptr::null());
        Self { raw }
    }

    /// Per-phase device times of the last call.
    pub fn timings(&self) -> ffi::pcuda_timings {
        let mut t = ffi::pcuda_timings::default();
        unsafe { ffi::pcuda_get_timings(self.raw, &mut t) };
        t
    }

    /// Multi-GPU, one process per GPU: rank 0 makes the id, the host program ships the 128 bytes to
    /// every rank (MPI, sockets …), every rank calls `comm_init`.
    pub fn comm_unique_id(&mut self) -> [u8; ffi::PCUDA_UNIQUE_ID_BYTES] {
        let mut id = [0u8; ffi::PCUDA_UNIQUE_ID_BYTES];
        check(unsafe { ffi::pcuda_comm_unique_id(self.raw, id.as_mut_ptr()) }, self.raw);
        id
    }

    pub fn comm_init(&mut self, id: &[u8; ffi::PCUDA_UNIQUE_ID_BYTES], world_size: i32, rank: i32) {
        check(unsafe { ffi::pcuda_comm_init(self.raw, id.as_ptr(), world_size, rank) }, self.raw);
    }

    pub fn raw(&mut self) -> *mut ffi::pcuda_ctx {
        self.raw
    }
}

impl Drop for CudaContext {
    fn drop(&mut self) {
        unsafe { ffi::pcuda_destroy(self.raw) }
    }
}

/// How an interaction maps onto the CUDA kernels — the analogue of `InteractionShader<P1, P2>`
/// (gpu/mod.rs:40-82): instead of WGSL source and buffer sizes it names the kernel family
/// (scalar, dimension), the softening and the `CHECKED` flag, and packs / unpacks the wire layout of
/// `include/particular_cuda.h` (`GravitationalField::from(&p)`, gravity/mod.rs:150-161).
pub trait CudaInteraction<P1, P2> {
    type Output;
    fn brute_force(&self, ctx: &mut CudaContext, affected: &[P1], affecting: &[P2]) -> Vec<Self::Output>;
    fn barnes_hut(&self, ctx: &mut CudaContext, theta: f64, affected: &[P1], affecting: &[P2]) -> Vec<Self::Output>;
}

macro_rules! impl_cuda_interaction {
    // $vec: glam vector, $s: scalar, $d: dimension, [$($c),*]: components,
    // $brute / $barnes: C entry points (barnes = None for f64)
    ($vec:ty, $s:ty, $d:literal, [$($c:ident),*], $brute:path, $barnes:expr) => {
        impl<const CHECKED: bool, P1, P2> CudaInteraction<P1, P2> for Acceleration<CHECKED>
        where
            P1: Position<Vector = $vec>,
            P2: Position<Vector = $vec> + Mass<Scalar = $s>,
        {
            type Output = $vec;
            fn brute_force(&self, ctx: &mut CudaContext, affected: &[P1], affecting: &[P2]) -> Vec<$vec> {
                run::<$vec, $s, $d, P1, P2>(ctx, affected, affecting, 0.0 as $s, CHECKED, None,
                    |p| { let v = p.position(); [$(v.$c),*] }, |p| { let v = p.position(); [$(v.$c),*] },
                    $brute, $barnes)
            }
            fn barnes_hut(&self, ctx: &mut CudaContext, theta: f64, affected: &[P1], affecting: &[P2]) -> Vec<$vec> {
                run::<$vec, $s, $d, P1, P2>(ctx, affected, affecting, 0.0 as $s, CHECKED, Some(theta),
                    |p| { let v = p.position(); [$(v.$c),*] }, |p| { let v = p.position(); [$(v.$c),*] },
                    $brute, $barnes)
            }
        }

        impl<const CHECKED: bool, P1, P2> CudaInteraction<P1, P2> for AccelerationSoftened<$s, CHECKED>
        where
            P1: Position<Vector = $vec>,
            P2: Position<Vector = $vec> + Mass<Scalar = $s>,
        {
            type Output = $vec;
            fn brute_force(&self, ctx: &mut CudaContext, affected: &[P1], affecting: &[P2]) -> Vec<$vec> {
                run::<$vec, $s, $d, P1, P2>(ctx, affected, affecting, self.softening, CHECKED, None,
                    |p| { let v = p.position(); [$(v.$c),*] }, |p| { let v = p.position(); [$(v.$c),*] },
                    $brute, $barnes)
            }
            fn barnes_hut(&self, ctx: &mut CudaContext, theta: f64, affected: &[P1], affecting: &[P2]) -> Vec<$vec> {
                run::<$vec, $s, $d, P1, P2>(ctx, affected, affecting, self.softening, CHECKED, Some(theta),
                    |p| { let v = p.position(); [$(v.$c),*] }, |p| { let v = p.position(); [$(v.$c),*] },
                    $brute, $barnes)
            }
        }
    };
}

type BruteFn<S> = unsafe extern "C" fn(*mut ffi::pcuda_ctx, *const S, usize, *const S, usize, S, i32, *mut S) -> i32;
type BarnesFn<S> = unsafe extern "C" fn(*mut ffi::pcuda_ctx, *const S, usize, *const S, usize, S, S, i32, *mut S) -> i32;

/// The two scalar types of the device kernels.
pub trait DeviceScalar: Copy + Default + Into<f64> {
    fn from_f64(v: f64) -> Self;
}
impl DeviceScalar for f32 {
    fn from_f64(v: f64) -> Self { v as f32 }
}
impl DeviceScalar for f64 {
    fn from_f64(v: f64) -> Self { v }
}

/// Pack, call, unpack.  `affected` and `affecting` being the same slice (the `&[P]` storage,
/// storage.rs:231-241) is detected by address and passed as `affected == NULL`, which lets the
/// targets alias the sources on the device and saves one upload.
#[allow(clippy::too_many_arguments)]
fn run<V, S, const D: usize, P1, P2>(
    ctx: &mut CudaContext, affected: &[P1], affecting: &[P2], softening: S, checked: bool,
    theta: Option<f64>, pos1: impl Fn(&P1) -> [S; D], pos2: impl Fn(&P2) -> [S; D],
    brute: BruteFn<S>, barnes: Option<BarnesFn<S>>,
) -> Vec<V>
where
    V: From<[S; D]>,
    S: DeviceScalar,
    P2: Mass<Scalar = S>,
{
    let mut src: Vec<S> = Vec::with_capacity(affecting.len() * (D + 1));
    for p in affecting {
        src.extend_from_slice(&pos2(p));
        src.push(p.mu());
    }
    let aliased = std::ptr::eq(affected.as_ptr().cast::<u8>(), affecting.as_ptr().cast::<u8>())
        && affected.len() == affecting.len()
        && std::mem::size_of::<P1>() == std::mem::size_of::<P2>();
    let mut tgt: Vec<S> = Vec::new();
    if !aliased {
        tgt.reserve(affected.len() * D);
        for p in affected {
            tgt.extend_from_slice(&pos1(p));
        }
    }
    let mut out = vec![[S::default(); D]; affected.len()];
    let tgt_ptr = if aliased { std::ptr::null() } else { tgt.as_ptr() };
    let status = match (theta, barnes) {
        (None, _) => unsafe {
            brute(ctx.raw, tgt_ptr, affected.len(), src.as_ptr(), affecting.len(), softening, checked as i32,
                  out.as_mut_ptr().cast())
        },
        (Some(theta), Some(barnes)) => unsafe {
            barnes(ctx.raw, tgt_ptr, affected.len(), src.as_ptr(), affecting.len(), S::from_f64(theta),
                   softening, checked as i32, out.as_mut_ptr().cast())
        },
        (Some(_), None) => unimplemented!("no device Barnes-Hut for this vector type"),
    };
    check(status, ctx.raw);
    out.into_iter().map(V::from).collect()
}

impl_cuda_interaction!(glam::Vec3, f32, 3, [x, y, z], ffi::pcuda_bruteforce_f32x3, Some(ffi::pcuda_barneshut_f32x3 as BarnesFn<f32>));
impl_cuda_interaction!(glam::Vec3A, f32, 3, [x, y, z], ffi::pcuda_bruteforce_f32x3, Some(ffi::pcuda_barneshut_f32x3 as BarnesFn<f32>));
impl_cuda_interaction!(glam::Vec2, f32, 2, [x, y], ffi::pcuda_bruteforce_f32x2, Some(ffi::pcuda_barneshut_f32x2 as BarnesFn<f32>));
impl_cuda_interaction!(glam::DVec3, f64, 3, [x, y, z], ffi::pcuda_bruteforce_f64x3, Some(ffi::pcuda_barneshut_f64x3 as BarnesFn<f64>));
impl_cuda_interaction!(glam::DVec2, f64, 2, [x, y], ffi::pcuda_bruteforce_f64x2, Some(ffi::pcuda_barneshut_f64x2 as BarnesFn<f64>));

// The reference's other vector front-ends (particular/Cargo.toml:19-34; their pair terms are wired
// at gravity/impls/nalgebra.rs and gravity/impls/ultraviolet.rs): same packing, same entry points.
// `v.x` reaches nalgebra's coordinates through Deref; both crates provide `From<[S; D]>`.
#[cfg(feature = "nalgebra")]
mod nalgebra_impls {
    use super::*;
    impl_cuda_interaction!(nalgebra::SVector<f32, 3>, f32, 3, [x, y, z], ffi::pcuda_bruteforce_f32x3, Some(ffi::pcuda_barneshut_f32x3 as BarnesFn<f32>));
    impl_cuda_interaction!(nalgebra::SVector<f32, 2>, f32, 2, [x, y], ffi::pcuda_bruteforce_f32x2, Some(ffi::pcuda_barneshut_f32x2 as BarnesFn<f32>));
    impl_cuda_interaction!(nalgebra::SVector<f64, 3>, f64, 3, [x, y, z], ffi::pcuda_bruteforce_f64x3, Some(ffi::pcuda_barneshut_f64x3 as BarnesFn<f64>));
    impl_cuda_interaction!(nalgebra::SVector<f64, 2>, f64, 2, [x, y], ffi::pcuda_bruteforce_f64x2, Some(ffi::pcuda_barneshut_f64x2 as BarnesFn<f64>));
}
#[cfg(feature = "ultraviolet")]
mod ultraviolet_impls {
    use super::*;
    impl_cuda_interaction!(ultraviolet::Vec3, f32, 3, [x, y, z], ffi::pcuda_bruteforce_f32x3, Some(ffi::pcuda_barneshut_f32x3 as BarnesFn<f32>));
    impl_cuda_interaction!(ultraviolet::Vec2, f32, 2, [x, y], ffi::pcuda_bruteforce_f32x2, Some(ffi::pcuda_barneshut_f32x2 as BarnesFn<f32>));
    impl_cuda_interaction!(ultraviolet::DVec3, f64, 3, [x, y, z], ffi::pcuda_bruteforce_f64x3, Some(ffi::pcuda_barneshut_f64x3 as BarnesFn<f64>));
    impl_cuda_interaction!(ultraviolet::DVec2, f64, 2, [x, y], ffi::pcuda_bruteforce_f64x2, Some(ffi::pcuda_barneshut_f64x2 as BarnesFn<f64>));
}

/// Brute-force algorithm on the GPU; same shape as `gpu::BruteForce<'a, T>` (gpu/mod.rs:149-177).
pub struct BruteForce<'a, T> {
    pub ctx: &'a mut CudaContext,
    pub interaction: T,
}

impl<'a, T> BruteForce<'a, T> {
    pub fn new(ctx: &'a mut CudaContext, interaction: T) -> Self {
        Self { ctx, interaction }
    }
}

impl<P1, P2, T> Interaction<Between<&[P1], &[P2]>> for BruteForce<'_, T>
where
    T: CudaInteraction<P1, P2>,
{
    type Output = std::vec::IntoIter<T::Output>; // as gpu::BruteForce, gpu/mod.rs:184

    fn compute(&mut self, Between(affected, affecting): Between<&[P1], &[P2]>) -> Self::Output {
        self.interaction.brute_force(self.ctx, affected, affecting).into_iter()
    }
}

/// Barnes-Hut on the GPU (sequential.rs:439-543 semantics: the tree is rebuilt on every call).
pub struct BarnesHut<'a, T> {
    pub ctx: &'a mut CudaContext,
    pub theta: f64,
    pub interaction: T,
}

impl<'a, T> BarnesHut<'a, T> {
    pub fn new(ctx: &'a mut CudaContext, theta: f64, interaction: T) -> Self {
        Self { ctx, theta, interaction }
    }
}

impl<P1, P2, T> Interaction<Between<&[P1], &[P2]>> for BarnesHut<'_, T>
where
    T: CudaInteraction<P1, P2>,
{
    type Output = std::vec::IntoIter<T::Output>;

    fn compute(&mut self, Between(affected, affecting): Between<&[P1], &[P2]>) -> Self::Output {
        self.interaction.barnes_hut(self.ctx, self.theta, affected, affecting).into_iter()
    }
}

/// Extension-trait sugar, the counterpart of `GpuCompute` (gpu/mod.rs:13-37).
pub trait CudaCompute<T>: Sized {
    fn cuda_brute_force<'a>(self, ctx: &'a mut CudaContext, interaction: T) -> <BruteForce<'a, T> as Interaction<Self>>::Output
    where
        BruteForce<'a, T>: Interaction<Self>,
    {
        BruteForce::new(ctx, interaction).compute(self)
    }

    fn cuda_barnes_hut<'a>(self, ctx: &'a mut CudaContext, theta: f64, interaction: T) -> <BarnesHut<'a, T> as Interaction<Self>>::Output
    where
        BarnesHut<'a, T>: Interaction<Self>,
    {
        BarnesHut::new(ctx, theta, interaction).compute(self)
    }
}

impl<T, P> CudaCompute<T> for &[P] {}
impl<T, P> CudaCompute<T> for &Ordered<P> {}
impl<T, P, F> CudaCompute<T> for &Reordered<'_, P, F> {}
impl<T, S1, S2> CudaCompute<T> for Between<S1, S2> {}

/// Which particles act on the others in a [`Simulation`].
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Affecting {
    All,
    /// `Reordered::new(particles, |p| p.mu() != 0)`: everything is affected, massive particles affect
    /// (storage.rs:153-163, 219-229; examples/ring-formation/src/nbody.rs:25-28).
    Massive,
}

/// Algorithm of a [`Simulation`].
#[derive(Clone, Copy, Debug)]
pub enum Algorithm {
    BruteForce,
    BarnesHut { theta: f64 },
}

/// Device-resident stepping: the loop every caller writes around `compute` —
/// `velocity += acceleration * dt; position += velocity * dt` (examples/simple/src/main.rs:45-59) —
/// with particles, velocities and accelerations kept in device memory between steps.  `S` is `f32`
/// (`D` = 2 or 3) or `f64` (`D` = 3, brute force).
pub struct Simulation<'a, S, const D: usize> {
    ctx: &'a mut CudaContext,
    raw: *mut ffi::pcuda_sim,
    n: usize,
    _scalar: PhantomData<S>,
}

pub trait SimScalar: Copy + Default {
    const TAG: u32;
}
impl SimScalar for f32 {
    const TAG: u32 = ffi::PCUDA_F32;
}
impl SimScalar for f64 {
    const TAG: u32 = ffi::PCUDA_F64;
}

impl<'a, S: SimScalar, const D: usize> Simulation<'a, S, D> {
    /// `particles`: `[x, y, (z,) mu]` rows — `GravitationalField` (gravity/mod.rs:12-18);
    /// `velocities`: one per particle or empty (at rest).
    #[allow(clippy::too_many_arguments)]
    pub fn new(ctx: &'a mut CudaContext, algorithm: Algorithm, softening: f64, checked: bool, dt: f64,
               affecting: Affecting, particles: &[([S; D], S)], velocities: &[[S; D]]) -> Self {
        assert!(velocities.is_empty() || velocities.len() == particles.len(), "one velocity per particle");
        let (alg, theta) = match algorithm {
            Algorithm::BruteForce => (ffi::PCUDA_BRUTE_FORCE, 0.0),
            Algorithm::BarnesHut { theta } => (ffi::PCUDA_BARNES_HUT, theta),
        };
        let cfg = ffi::pcuda_sim_config {
            dim: D as u32, scalar: S::TAG, algorithm: alg,
            flags: if affecting == Affecting::Massive { ffi::PCUDA_SIM_AFFECTING_MASSIVE_ONLY } else { 0 },
            theta, softening, dt, checked: checked as i32, reserved: 0,
        };
        let mut flat: Vec<S> = Vec::with_capacity(particles.len() * (D + 1));
        for (pos, mu) in particles {
            flat.extend_from_slice(pos);
            flat.push(*mu);
        }
        let vel_ptr = if velocities.is_empty() { std::ptr::null() } else { velocities.as_ptr().cast() };
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::pcuda_sim_create(ctx.raw, &cfg, flat.as_ptr().cast(), vel_ptr, particles.len(), &mut raw) }, ctx.raw);
        Self { ctx, raw, n: particles.len(), _scalar: PhantomData }
    }

    /// Enqueues `n_steps` steps; returns without waiting for the device.
    pub fn step(&mut self, n_steps: u32) {
        check(unsafe { ffi::pcuda_sim_step(self.ctx.raw, self.raw, n_steps) }, self.ctx.raw);
    }

    /// Blocking read-back of `[position.., mu]` rows.
    pub fn particles(&mut self) -> Vec<([S; D], S)> {
        let mut flat = vec![S::default(); self.n * (D + 1)];
        check(unsafe { ffi::pcuda_sim_read(self.ctx.raw, self.raw, flat.as_mut_ptr().cast(), std::ptr::null_mut(), std::ptr::null_mut()) }, self.ctx.raw);
        flat.chunks_exact(D + 1).map(|r| { let mut p = [S::default(); D]; p.copy_from_slice(&r[..D]); (p, r[D]) }).collect()
    }

    pub fn velocities(&mut self) -> Vec<[S; D]> {
        let mut out = vec![[S::default(); D]; self.n];
        check(unsafe { ffi::pcuda_sim_read(self.ctx.raw, self.raw, std::ptr::null_mut(), out.as_mut_ptr().cast(), std::ptr::null_mut()) }, self.ctx.raw);
        out
    }

    pub fn accelerations(&mut self) -> Vec<[S; D]> {
        let mut out = vec![[S::default(); D]; self.n];
        check(unsafe { ffi::pcuda_sim_read(self.ctx.raw, self.raw, std::ptr::null_mut(), std::ptr::null_mut(), out.as_mut_ptr().cast()) }, self.ctx.raw);
        out
    }
}

impl<S, const D: usize> Drop for Simulation<'_, S, D> {
    fn drop(&mut self) {
        unsafe { ffi::pcuda_sim_destroy(self.ctx.raw, self.raw) }
    }
}

/// A user-defined pair interaction compiled for the device at run time — the counterpart of
/// implementing `InteractionShader<P1, P2>` (gpu/mod.rs:40-82).  `source` is CUDA C++ defining
/// `struct Affected / Affecting / Interaction / Push` and
/// `__device__ void compute(const Affected&, const Affecting&, Interaction&)`.  `A`, `B`, `I`, `P` are
/// `#[repr(C)]` `Copy` types with the same layouts (checked against the device `sizeof`).
pub struct CustomInteraction<A, B, I, P = ()> {
    raw: *mut ffi::pcuda_interaction,
    _types: PhantomData<(A, B, I, P)>,
}

impl<A: Copy, B: Copy, I: Copy + Default, P: Copy> CustomInteraction<A, B, I, P> {
    pub fn new(ctx: &mut CudaContext, source: &str) -> Self {
        let src = std::ffi::CString::new(source).expect("interaction source contains a NUL byte");
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::pcuda_interaction_create(ctx.raw, src.as_ptr(), &mut raw) }, ctx.raw);
        let mut sizes = [0u32; 4];
        unsafe { ffi::pcuda_interaction_sizes(raw, sizes.as_mut_ptr()) };
        assert_eq!(
            [sizes[0] as usize, sizes[1] as usize, sizes[2] as usize],
            [std::mem::size_of::<A>(), std::mem::size_of::<B>(), std::mem::size_of::<I>()],
            "host types do not match the device structs"
        );
        Self { raw, _types: PhantomData }
    }

    /// `interactions[i] = fold over affecting of compute(affected[i], affecting[j], out)`.
    pub fn brute_force(&mut self, ctx: &mut CudaContext, affected: &[A], affecting: &[B], push: &P) -> Vec<I> {
        let mut out = vec![I::default(); affected.len()];
        check(unsafe {
            ffi::pcuda_interaction_brute_force(ctx.raw, self.raw, affected.as_ptr().cast(), affected.len(),
                affecting.as_ptr().cast(), affecting.len(), (push as *const P).cast(), std::mem::size_of::<P>(),
                out.as_mut_ptr().cast())
        }, ctx.raw);
        out
    }

    /// Frees the device module; must be called with the context that created it.
    pub fn destroy(self, ctx: &mut CudaContext) {
        unsafe { ffi::pcuda_interaction_destroy(ctx.raw, self.raw) }
    }
}

#[cfg(test)]
mod tests {
    //! The reference's own algorithm tests, instantiated for the CUDA operators: `tests_algorithms!`
    //! only needs `$cm.compute(&reordered)` / `$cm.compute(slice)` (gravity/newtonian/mod.rs:247, 319).
    use super::*;
    use particular::{acceleration_error, circular_orbit};

    #[test]
    fn brute_force_f32x3() {
        let mut ctx = CudaContext::new(0);
        acceleration_error!(BruteForce::new(&mut ctx, Acceleration::checked()), 1e-2, glam::Vec3, f32);
    }

    #[test]
    fn barnes_hut_05_f32x3() {
        let mut ctx = CudaContext::new(0);
        acceleration_error!(BarnesHut::new(&mut ctx, 0.5, Acceleration::checked()), 5e-1, glam::Vec3, f32);
    }

    #[test]
    fn circular_orbit_f32x3() {
        let mut ctx = CudaContext::new(0);
        circular_orbit!(BruteForce::new(&mut ctx, Acceleration::checked()), 100, 1e-2, glam::Vec3, f32);
    }
}
