"""GPU tuning run for Barnes-Hut: grouping threshold, leaf size, error statistics.
Usage (on the GPU box): python scripts/tune_bh.py [N]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle
import particular_b200 as pb
from particular_b200._ffi import lib
from tests.conftest import plummer_cloud, rel_err, uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
res = {}

# ---- 2-D brute-force tolerance diagnostic -------------------------------------------------------
ctx = pb.CudaContext(0)
for dim, n in ((2, 3001), (3, 3001), (2, 6000), (3, 16384)):
    p = uniform_cloud(n, d=dim, seed=7)
    exact = oracle.brute_force_exact(p[:, :dim], p)
    S = oracle.brute_force_abs(p[:, :dim], p)
    ref = oracle.brute_force_parallel(p[:, :dim], p)
    got = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(p)
    a = np.linalg.norm(exact, axis=1)
    kappa = S / a
    e_gpu, e_ref = rel_err(got, exact), rel_err(ref, exact)
    print(f"bf dim={dim} n={n}: kappa median {np.median(kappa):.1f} max {kappa.max():.1f}; "
          f"e_gpu max {e_gpu.max():.2e} e_ref max {e_ref.max():.2e} gpu-vs-ref max {rel_err(got, ref).max():.2e}; "
          f"normalised: gpu {np.max(e_gpu / kappa):.2e} ref {np.max(e_ref / kappa):.2e} "
          f"gpu-vs-ref {np.max(rel_err(got, ref) / kappa):.2e}")

# ---- Barnes-Hut ------------------------------------------------------------------------------------
P = plummer_cloud(N)
d_src = torch.from_numpy(P).cuda()
d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
idx = np.sort(np.random.default_rng(1).choice(N, 2048, replace=False))
exact = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(pb.Between(P[idx, :3], P))
ctx.close()
for leaf in (8, 16):
    ctx = pb.CudaContext(0, leaf_size=leaf)
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    for seg in (64, 128, 256):
        lib.pcuda_debug_set(b"bh_seg_max", seg)
        for count in (1, 0):
            lib.pcuda_debug_set(b"bh_count", count)
            for _ in range(2):
                bh.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
                ctx.sync()
            t = ctx.timings()
            if count:
                c = bh.last_counters()
                err = rel_err(d_out.cpu().numpy()[idx], exact)
                print(f"leaf={leaf} seg={seg}: build {t['build_ms']:.2f} ms traverse {t['compute_ms']:.2f} ms "
                      f"launches {t['kernel_launches']}; per target: node {c['node_interactions']/N:.0f} "
                      f"particle {c['particle_interactions']/N:.0f}; tests {c['node_tests']:.3e}; "
                      f"err median {np.median(err):.2e} p99 {np.percentile(err, 99):.2e} max {err.max():.2e}")
                res[f"leaf{leaf}_seg{seg}"] = {"build_ms": t["build_ms"], "traverse_ms": t["compute_ms"], **c}
            else:
                print(f"    (counters off) traverse {t['compute_ms']:.2f} ms")
    ctx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/tune_bh.json", "w"))
