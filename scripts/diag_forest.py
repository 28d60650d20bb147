"""Accuracy of the key-range-partitioned Barnes-Hut build (virtual ranks on one GPU) against the single tree and the
exact sum at small N, where the cells next to the cuts are a large share of the tree.  Usage: python scripts/diag_forest.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
from tests.conftest import uniform_cloud, plummer_cloud
def st(e): return f"med {np.median(e):.3e} p90 {np.percentile(e,90):.3e} p99 {np.percentile(e,99):.3e} max {e.max():.3e}"
with pb.CudaContext(0) as ctx:
    for name, p in (("uniform20k", uniform_cloud(20000, seed=5)), ("plummer20k", plummer_cloud(20000, seed=5)), ("uniform200k", uniform_cloud(200000, seed=5))):
        bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
        exact = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p)
        den = np.linalg.norm(exact, axis=1)
        one = bh.compute(p)
        print(name, "single:", st(np.linalg.norm(one - exact, axis=1) / den), bh.last_counters())
        for parts in (2, 3, 8, 16):
            j = bh.compute_partitioned(p, parts)
            e = np.linalg.norm(j - exact, axis=1) / den
            d = np.linalg.norm(j - one, axis=1) / den
            print(f"  parts {parts:2d}: err {st(e)} | diff vs single {st(d)} frac>1e-4 {np.mean(d > 1e-4):.3f}", bh.last_counters())
