//! 0.7-era names on top of the 0.8 traits (particular/CHANGELOG.md:8-35): what BASELINE.json's
//! north_star calls `ComputeMethod`, `#[derive(Particle)]` and `.accelerations(&mut method)`.
use particular::gravity::{Mass, Position};
use particular::Interaction;

/// 0.7 `ComputeMethod<S>` == 0.8 `Interaction<S>` (CHANGELOG.md:23).
pub trait ComputeMethod<S>: Interaction<S> {}
impl<S, C: Interaction<S>> ComputeMethod<S> for C {}

/// 0.7 `Particle` == `Position + Mass` (CHANGELOG.md:31); `#[derive(Position, Mass)]` provides it.
pub trait Particle: Position + Mass {}
impl<P: Position + Mass> Particle for P {}

/// The removed `.accelerations(&mut compute_method)` adaptor (CHANGELOG.md:33):
/// `bodies.accelerations(&mut cuda::BruteForce::new(&mut ctx, Acceleration::checked()))`.
pub trait Accelerations<'p, P: 'p>: Sized {
    fn accelerations<C: Interaction<&'p [P]>>(self, compute_method: &mut C) -> C::Output;
}

impl<'p, P> Accelerations<'p, P> for &'p [P] {
    fn accelerations<C: Interaction<&'p [P]>>(self, compute_method: &mut C) -> C::Output {
        compute_method.compute(self)
    }
}

/// 0.7 type names of the CUDA operators.
pub type BruteForceCuda<'a, T> = crate::BruteForce<'a, T>;
pub type BarnesHutCuda<'a, T> = crate::BarnesHut<'a, T>;
