"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck /
synccheck).  Usage (GPU box): compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
from tests.conftest import plummer_cloud, uniform_cloud

with pb.CudaContext(0) as ctx:
    for n in (1, 37, 700, 5000):                      # brute force: direct, fused-ticket and split paths
        p = uniform_cloud(n, seed=n)
        pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p)
        pb.BruteForce(ctx, pb.AccelerationSoftened.checked(3.0)).compute(pb.Reordered(p))
    p2 = uniform_cloud(900, d=2, seed=5)
    pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p2)
    p64 = uniform_cloud(600, dtype=np.float64, seed=6)
    pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p64)
    big = uniform_cloud(40000, seed=8)                # large enough for several source splits + reduce kernel
    pb.BruteForce(ctx, pb.Acceleration.checked()).compute(pb.Between(big[:3000, :3], big))
    for n, d in ((1, 3), (50, 3), (3000, 3), (3000, 2), (40000, 3)):   # build_small and per-level build
        q = plummer_cloud(n, d=d, seed=n)
        pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(q)
    for n, d in ((1, 3), (300, 2), (3000, 3), (40000, 3)):   # f64 layer: gather64, moments64, traverse64
        q64 = plummer_cloud(n, d=d, seed=n, dtype=np.float64)
        pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(q64)
        pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(0.1)).compute(pb.Between(q64[: n // 2 + 1, :d] * 0.5, q64))
    pb.BruteForce(ctx, pb.Acceleration.checked()).compute(uniform_cloud(700, d=2, dtype=np.float64, seed=6))
    pb.morton_keys(ctx, plummer_cloud(5000, seed=4))
    sb = pb.ShardedBetween(ctx, pb.AccelerationSoftened.checked(1.0))   # single rank: no communicator
    sb.compute(pb.Reordered.new(uniform_cloud(3000, seed=7, massive_ratio=0.05)))
    q = plummer_cloud(40000, seed=3)
    tree = pb.RootedOrthtree(ctx, q)
    pb.BarnesHut(ctx, 0.7, pb.Acceleration.checked()).compute(pb.Between(uniform_cloud(777, seed=1)[:, :3] * 1e-3, tree))
    tree.close()
    with pb.Simulation(pb.BruteForce(ctx, pb.Acceleration.checked()), uniform_cloud(500, seed=2), dt=1e-3,
                       affecting="massive") as sim:
        sim.step(19)
        sim.read(True, True, True)
    with pb.Simulation(pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()), plummer_cloud(2000, seed=2), dt=1e-3) as sim:
        sim.step(2)
        sim.particles()
print("sanitize_smoke: done")
