"""particular_b200 — B200 (sm_100a) compute backend for the `particular` N-body crate.

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, libparticular_cuda.so) and the
host-side mirror of the reference's operator interface (interface.py).  See DESIGN.md.

The interface is loaded on first attribute access so that ``python -m particular_b200.build`` can
run before the library exists; any use of the API without the built library raises ImportError
(there is no CPU fallback).
"""
__all__ = ["Acceleration", "AccelerationSoftened", "BarnesHut", "Between", "BruteForce",
           "CudaContext", "CudaError", "CustomInteraction", "check_interaction_source", "Ordered", "Reordered", "RootedOrthtree", "Simulation",
           "cuda_barnes_hut", "cuda_brute_force", "is_affecting", "ShardedBruteForce",
           "ShardedBarnesHut", "ShardedBetween", "morton_keys",
           "shard_bounds", "shard_capacity"]


def __getattr__(name):
    if name in ("ShardedBruteForce", "ShardedBarnesHut", "ShardedBetween", "shard_bounds",
                "shard_capacity"):
        from . import sharded
        return getattr(sharded, name)
    if name in __all__:
        from . import interface
        return getattr(interface, name)
    raise AttributeError(f"module 'particular_b200' has no attribute {name!r}")
