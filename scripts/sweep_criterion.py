"""The reference's criterion benchmark shape on the CUDA backend (benches/benchmark.rs: N = 2^1..2^16,
3-D f32, all bodies massive, uniform cube, seed 1808, Acceleration::checked(), theta in {0.3, 0.7};
every sample = host slice in -> host result out, i.e. upload + kernels + read-back, like the
reference's gpu::BruteForce rows).  Prints one JSON line per (function id, N) with the mean time in
ns, and the published reference numbers beside it where one exists.
Usage (on the GPU box): python scripts/sweep_criterion.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
from tests.conftest import uniform_cloud

PUBLISHED_NS = {  # particular/benches/results/native-x86-f32-3d.json (Ryzen 7900X / RTX 3080)
    ("gpu::BruteForce", 1024): 0.140e6, ("gpu::BruteForce", 16384): 0.542e6,
    ("gpu::BruteForce", 65536): 5.886e6, ("parallel::BruteForceSIMD<8>", 1024): 0.0678e6,
    ("parallel::BruteForceSIMD<8>", 65536): 76.92e6, ("sequential::BruteForceScalar", 1024): 1.777e6,
    ("parallel::BarnesHut::0.3", 65536): 85.17e6, ("parallel::BarnesHut::0.7", 65536): 20.21e6,
}
COMPARE = {"cuda::BruteForce": ["gpu::BruteForce", "parallel::BruteForceSIMD<8>", "sequential::BruteForceScalar"],
           "cuda::BarnesHut::0.3": ["parallel::BarnesHut::0.3"], "cuda::BarnesHut::0.7": ["parallel::BarnesHut::0.7"]}


def measure(fn, min_time=0.3, min_samples=15):
    for _ in range(3):
        fn()
    times = []
    t_end = time.perf_counter() + min_time
    while len(times) < min_samples or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
        if len(times) >= 2000:
            break
    return 1e9 * float(np.mean(times)), 1e9 * float(np.median(times)), len(times)


ctx = pb.CudaContext(0, phase_timings=False)  # the lean product path: no per-phase events
rows = []
for k in range(1, 17):
    n = 2 ** k
    p = ctx.pinned_empty((n, 4), np.float32)
    p[:] = uniform_cloud(n, seed=1808)
    out = ctx.pinned_empty((n, 3), np.float32)
    algos = {"cuda::BruteForce": pb.BruteForce(ctx, pb.Acceleration.checked()),
             "cuda::BarnesHut::0.3": pb.BarnesHut(ctx, 0.3, pb.Acceleration.checked()),
             "cuda::BarnesHut::0.7": pb.BarnesHut(ctx, 0.7, pb.Acceleration.checked())}
    for fid, algo in algos.items():
        mean, med, cnt = measure(lambda: algo.compute(p, out=out))
        row = {"id": f"Particular/{fid}/{n}", "n": n, "mean_ns": mean, "median_ns": med, "samples": cnt}
        for ref in COMPARE[fid]:
            if (ref, n) in PUBLISHED_NS:
                row[f"published {ref} ns"] = PUBLISHED_NS[(ref, n)]
        rows.append(row)
        print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_criterion.json", "w"), indent=1)
