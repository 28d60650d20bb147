"""The key-range-partitioned Barnes-Hut build (virtual ranks on one GPU) for compute-sanitizer:
partition kernels, per-part builds (single-launch and per-level), pack / boundary collection, top tree,
forest walk.  Usage (GPU box): compute-sanitizer --tool memcheck python scripts/sanitize_partitioned.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
from tests.conftest import plummer_cloud, uniform_cloud

with pb.CudaContext(0) as ctx:
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    for n, parts in ((3, 8), (700, 2), (5000, 3), (20000, 16), (150000, 4)):   # 150000 / 4: per-level build
        q = plummer_cloud(n, seed=n)
        out = bh.compute_partitioned(q, parts)
        assert np.isfinite(out).all()
    same = np.tile(np.array([[1.0, 2.0, 3.0, 5.0]], np.float32), (300, 1))
    pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(0.5)).compute_partitioned(same, 4)
    pb.BarnesHut(ctx, 0.0, pb.Acceleration.checked()).compute_partitioned(uniform_cloud(2000, seed=1), 5)
print("sanitize_partitioned: done")
