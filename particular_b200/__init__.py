"""particular_b200 — B200 (sm_100a) compute backend for the `particular` N-body crate.

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, libparticular_cuda.so) and the
host-side mirror of the reference's operator interface (interface.py).  See DESIGN.md.
"""
from .interface import (Acceleration, AccelerationSoftened, BarnesHut, Between, BruteForce,
                        CudaContext, CudaError, Ordered, Reordered, RootedOrthtree,
                        cuda_barnes_hut, cuda_brute_force, is_affecting)

__all__ = ["Acceleration", "AccelerationSoftened", "BarnesHut", "Between", "BruteForce",
           "CudaContext", "CudaError", "Ordered", "Reordered", "RootedOrthtree",
           "cuda_barnes_hut", "cuda_brute_force", "is_affecting"]
