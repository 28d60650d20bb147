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
    return 1e9 * np.asarray(times)


def estimate(x, stat, rng, resamples=200):
    """criterion-style estimate: point estimate + bootstrap 95 % interval and standard error."""
    point = float(stat(x))
    boots = np.array([stat(x[rng.integers(0, len(x), len(x))]) for _ in range(resamples)])
    lo, hi = np.percentile(boots, [2.5, 97.5])
    return {"confidence_interval": {"confidence_level": 0.95, "lower_bound": float(lo),
                                    "upper_bound": float(hi)},
            "point_estimate": point, "standard_error": float(boots.std())}


def criterion_entry(name, function_id, n, times_ns, rng):
    """One entry in the schema of the reference's benches/results/*.json (what its benchmark site
    reads): criterion_benchmark_v1 + criterion_estimates_v1.  Every sample is one iteration, so
    `slope` (criterion's per-iteration regression estimate) is reported as the mean."""
    full_id = f"Particular/{function_id}/{n}"
    mad = lambda v: np.median(np.abs(v - np.median(v)))  # noqa: E731
    return full_id, {
        "baseline": name, "fullname": f"{name}/{full_id}",
        "criterion_benchmark_v1": {
            "group_id": "Particular", "function_id": function_id, "value_str": str(n),
            "throughput": None, "full_id": full_id,
            "directory_name": f"particular/{function_id.lower().replace('::', '__')}/{n}"},
        "criterion_estimates_v1": {
            "mean": estimate(times_ns, np.mean, rng), "median": estimate(times_ns, np.median, rng),
            "median_abs_dev": estimate(times_ns, mad, rng), "slope": estimate(times_ns, np.mean, rng),
            "std_dev": estimate(times_ns, np.std, rng)}}


EXPORT_NAME = "native-b200-f32-3d"
export = {"name": EXPORT_NAME, "tags": ["native", "cuda", "sm_100a", "f32", "3d", "B200"], "benchmarks": {}}
boot_rng = np.random.default_rng(1808)
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
        t_ns = measure(lambda: algo.compute(p, out=out))
        mean, med, cnt = float(t_ns.mean()), float(np.median(t_ns)), len(t_ns)
        key, entry = criterion_entry(EXPORT_NAME, fid, n, t_ns, boot_rng)
        export["benchmarks"][key] = entry
        row = {"id": f"Particular/{fid}/{n}", "n": n, "mean_ns": mean, "median_ns": med, "samples": cnt}
        for ref in COMPARE[fid]:
            if (ref, n) in PUBLISHED_NS:
                row[f"published {ref} ns"] = PUBLISHED_NS[(ref, n)]
        rows.append(row)
        print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_criterion.json", "w"), indent=1)
json.dump(export, open(f"gpurun_out/{EXPORT_NAME}.json", "w"), indent=4)
