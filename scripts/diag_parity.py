"""Brute-force parity diagnostic: GPU vs the restated f32 fold and vs the exact sum, by condition-number
class (input to tests/conftest.py: parity_tolerance).  Usage (GPU box): python scripts/diag_parity.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import particular_b200 as pb
from tests.conftest import EPS32, EPS64, rel_err, uniform_cloud

ctx = pb.CudaContext(0)
for dtype, u in ((np.float32, EPS32), (np.float64, EPS64)):
    for dim in (3, 2):
        for n in (1000, 3001, 6000, 16384):
            p = uniform_cloud(n, d=dim, seed=7, dtype=dtype)
            exact = oracle.brute_force_exact(p[:, :dim], p)
            S = oracle.brute_force_abs(p[:, :dim], p)
            ref = oracle.brute_force_parallel(p[:, :dim], p)
            got = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p)
            kappa = S / np.linalg.norm(exact, axis=1)
            e = rel_err(got, ref)
            base = 1e-5 if dtype == np.float32 else 1e-12
            cut = base / (np.sqrt(n) * u)
            line = f"{np.dtype(dtype).name} dim={dim} n={n}: kappa_cut {cut:.2f}"
            for lo, hi in ((0, cut), (cut, 4 * cut), (4 * cut, 1e30)):
                m = (kappa > lo) & (kappa <= hi)
                if m.any():
                    line += (f" | kappa in ({lo:.1f},{hi:.1f}]: {m.sum()} particles, max gpu-vs-fold {e[m].max():.2e}, "
                             f"fold-vs-exact {rel_err(ref, exact)[m].max():.2e}, gpu-vs-exact {rel_err(got, exact)[m].max():.2e}")
            print(line, flush=True)
ctx.close()
