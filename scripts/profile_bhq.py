"""Quadrupole knob: time and error of Barnes-Hut with expansion_order 1 and 2 over theta.
Usage (GPU box): python scripts/profile_bhq.py [N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from tests.conftest import plummer_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
P = plummer_cloud(N)
d_src = torch.from_numpy(P).cuda()
d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
idx = np.linspace(0, N - 1, 4096).astype(np.int64)
with pb.CudaContext(0) as c1:
    exact = pb.BruteForce(c1, pb.Acceleration.checked()).compute(pb.Between(P[idx, :3].astype(np.float64), P.astype(np.float64)))
for order in (1, 2):
    with pb.CudaContext(0, expansion_order=order) as ctx:
        for theta in (0.3, 0.5, 0.7, 0.9, 1.1):
            bh = pb.BarnesHut(ctx, theta, pb.Acceleration.checked())
            for _ in range(3):
                bh.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
                ctx.sync()
            t = ctx.timings()
            a = d_out.cpu().numpy()[idx].astype(np.float64)
            e = np.linalg.norm(a - exact, axis=1) / np.linalg.norm(exact, axis=1)
            print(f"order {order} theta {theta}: build {t['build_ms']:.3f} ms traverse {t['compute_ms']:.3f} ms "
                  f"err median {np.median(e):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e}", flush=True)
