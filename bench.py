#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 backend for particular's hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload bruteforce|barneshut] [--n N] [--no-extra]

Default workload (BASELINE.json configs[1]): brute force, 3-D f32, N = 1,000,000 massive particles
drawn like the reference's criterion bench (benches/benchmark.rs:22-37: positions U[-5e3,5e3)^3,
mu U[1e3,1e9), seed 1808), `Acceleration::checked()`; one "step" = one full evaluation of all N x N
pair interactions (self pairs counted, as the reference evaluates them).  Metric: Gpair-interactions/s.

  value   device-resident: the particles live in HBM; every step = pad/copy into the gather slot,
          (N > 1 GPUs) in-place NCCL all-gather of the source records, pair kernel, fixed-order
          reduction of the source-split partial sums.  N > 1: strong scaling — the N particles are
          sharded over the ranks (targets), sources replicated by the all-gather.
  e2e     the same evaluation through the public API with HOST buffers (pinned upload of the
          rank's records, step, download of its accelerations inside the timed region).
  roofline  FP32 pipe: 20 flop / pair (the GPU-Gems-3 convention the reference cites,
          gpu/resources.rs:76-77) over the pair kernel's CUDA-event time.
  cpu_baseline  the restated `parallel::BruteForceSimd<8>` (oracle/baseline_simd.c; AVX2 + OpenMP)
          on the box's host cores, on a bounded target sample of the same workload.

`--impl reference` times that CPU restatement as the reference arm (the Rust crate cannot be built
here: no cargo/rustc in the image — DESIGN.md).  `--workload barneshut` makes Barnes-Hut
(theta = 0.5, Plummer sphere, N = 10M) the headline line instead; by default its number rides along
in the "barnes_hut" key of the brute-force line at N = 1 GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 20.0
SEED = 1808


# ---- synthetic workloads -----------------------------------------------------------
def uniform_cloud(n, seed=SEED):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), dtype=np.float32)
    p[:, :3] = rng.uniform(-5e3, 5e3, (n, 3))
    p[:, 3] = rng.uniform(1e3, 1e9, n)
    return p


def plummer_cloud(n, seed=SEED):
    rng = np.random.default_rng(seed)
    r = np.empty(0)
    while len(r) < n:
        u = rng.uniform(1e-12, 1.0, n)
        rr = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr < 50.0]])
    r = r[:n]
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    p = np.empty((n, 4), dtype=np.float32)
    p[:, :3] = v * r[:, None]
    p[:, 3] = 1.0 / n
    return p


# ---- clocks during the timed region (NVML; B200_PROFILING.md "clocks line") -------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report nothing rather than invent numbers
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": self.err}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "sm_mhz_min": s[0] if s else None, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(s)}


def measured_hbm_peak():
    """HBM copy bandwidth measured by the driver on this pool (MEASURED_PEAKS.json), else the
    profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by scripts/ncu_digest.py --traffic; measured under the
    profiler at the bench's own workload, so it is reported beside the timing, never derived
    from it).  None when no capture is committed for this kernel."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[kernel_key]
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), t.get("source")
    except Exception:
        return None, None


# ---- the other BASELINE.json configs + the stepping path (ride along at 1 GPU) -----------------------
def other_configs(ctx, stream, flush, steps=3):
    """configs[2] (10k massive + 16M massless split), configs[4] (f64 N=256k; 2-D Barnes-Hut N=4M)
    and device-resident stepping at the reference's criterion size: device-resident times with
    CUDA events on the context stream, L2 flushed between steps."""
    import torch

    import particular_b200 as pb
    dev = torch.device("cuda", ctx.device)
    out = {}

    def timed(fn, k=steps, warm=2):
        for _ in range(warm):
            fn()
        ctx.sync()
        ts = []
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.sync()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    rng = np.random.default_rng(SEED)
    # configs[2]: ring-formation style split, `Reordered`: all particles affected, massive affecting
    n_massive, n_massless = 10_000, 16_000_000
    src = uniform_cloud(n_massive)
    d_src = torch.from_numpy(src).to(dev)
    tgt = np.empty((n_massive + n_massless, 3), np.float32)
    tgt[:n_massive] = src[:, :3]
    tgt[n_massive:] = rng.uniform(-5e3, 5e3, (n_massless, 3))
    d_tgt = torch.from_numpy(tgt).to(dev)
    d_out = torch.empty_like(d_tgt)
    bf = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(1.0))
    ms = timed(lambda: bf.compute_device(d_tgt.data_ptr(), len(tgt), d_src.data_ptr(), n_massive,
                                         d_out.data_ptr()))
    h_tgt = ctx.pinned_empty(tgt.shape, np.float32)
    h_tgt[:] = tgt
    h_out = ctx.pinned_empty(tgt.shape, np.float32)
    e2e = timed(lambda: bf.compute(pb.Between(h_tgt, src), out=h_out), k=2, warm=1)
    pairs = float(len(tgt)) * n_massive
    out["split_10k_massive_16M_massless"] = {
        "config": "BASELINE configs[2]: 10,000 massive + 16,000,000 massless, Between(all, massive) "
                  "(Reordered storage), AccelerationSoftened::checked(1.0)",
        "pairs_per_step": pairs, "ms_per_step": ms, "value": pairs / ms / 1e6, "unit": "Gpairs/s",
        "e2e": {"ms_per_step": e2e, "value": pairs / e2e / 1e6, "unit": "Gpairs/s",
                "h2d_bytes_per_step": int(tgt.nbytes + src.nbytes), "d2h_bytes_per_step": int(tgt.nbytes)}}
    del d_tgt, d_out, d_src

    # configs[4a]: f64 precision path
    n64 = 262_144
    p64 = uniform_cloud(n64).astype(np.float64)
    d_p = torch.from_numpy(p64).to(dev)
    d_o = torch.empty((n64, 3), dtype=torch.float64, device=dev)
    bf64 = pb.BruteForce(ctx, pb.Acceleration.checked())
    ms = timed(lambda: bf64.compute_device(None, n64, d_p.data_ptr(), n64, d_o.data_ptr(), "f64x3"))
    peak64 = ctx.sm_count * 64 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    tf = FLOP_PER_PAIR * float(n64) * n64 / (ms * 1e-3) / 1e12
    out["f64_256k"] = {"config": "BASELINE configs[4]: brute force 3-D f64, N=262144, Acceleration::checked()",
                       "ms_per_step": ms, "value": float(n64) * n64 / ms / 1e6, "unit": "Gpairs/s",
                       "fp64_tflops_20flop_per_pair": tf, "fp64_peak_tflops_nominal": peak64,
                       "frac": tf / peak64}
    del d_p, d_o

    # configs[4b]: particle-toy style 2-D quadtree
    n2 = 4_194_304
    p2 = np.empty((n2, 3), np.float32)
    p2[:, :2] = rng.uniform(-5e3, 5e3, (n2, 2))
    p2[:, 2] = rng.uniform(1e3, 1e9, n2)
    d_p = torch.from_numpy(p2).to(dev)
    d_o = torch.empty((n2, 2), dtype=torch.float32, device=dev)
    bh2 = pb.BarnesHut(ctx, 0.5, pb.AccelerationSoftened.checked(100.0))
    ms = timed(lambda: bh2.compute_device(None, n2, d_p.data_ptr(), n2, d_o.data_ptr(), "f32x2"))
    t = ctx.timings()
    out["barnes_hut_2d_4M"] = {"config": "BASELINE configs[4]: Barnes-Hut 2-D f32 quadtree, theta=0.5, N=4194304 "
                                         "uniform square, AccelerationSoftened::checked(100)",
                               "ms_per_step": ms, "build_ms": t["build_ms"], "traverse_ms": t["compute_ms"],
                               "value": n2 / (ms * 1e-3), "unit": "particles/s"}
    del d_p, d_o

    # device-resident stepping at the reference's criterion size (benches/benchmark.rs: N = 2^k)
    nb = 1024
    pb_small = uniform_cloud(nb)
    inter = pb.Acceleration.checked()
    res = {}
    for label, graph in (("graph", True), ("eager", False)):
        with pb.Simulation(pb.BruteForce(ctx, inter), pb_small, dt=1e-3, graph=graph) as sim:
            sim.step(64)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            sim.step(2048)
            e1.record(stream)
            ctx.sync()
            wall = time.perf_counter() - t0
            res[label] = {"us_per_step_device": 1e3 * e0.elapsed_time(e1) / 2048,
                          "us_per_step_wall": 1e6 * wall / 2048}
    one = pb.BruteForce(ctx, inter)
    one.compute(pb_small)
    t0 = time.perf_counter()
    for _ in range(200):
        one.compute(pb_small)
    res["one_shot_host_api"] = {"us_per_step_wall": 1e6 * (time.perf_counter() - t0) / 200}
    out["stepping_1024"] = {"config": "device-resident stepping (pcuda_sim_*), brute force 3-D f32, N=1024 "
                                      "(criterion bench shape), 2048 steps; one_shot_host_api = upload + "
                                      "kernel + read-back per step, as the reference's wgpu operator works",
                            **res}
    return out


# ---- CPU arms ----------------------------------------------------------------------------------------
def cpu_bruteforce_rate(P, seconds, steps=1, targets=None, softening=0.0):
    """Restated parallel::BruteForceSimd<8> on a bounded target sample x all sources.
    `targets`: affected positions when they are not the sources themselves.
    Returns (Gpairs/s, sample description, cores, per-step seconds list)."""
    import oracle
    n = len(P)
    T = P[:, :3] if targets is None else targets
    nt = len(T)
    oracle.use_all_cores()
    cores = oracle.baseline_threads()
    probe = min(nt, 256 * cores)
    t0 = time.perf_counter()
    oracle.brute_force_simd8_parallel(T[:probe], P, softening)
    dt = max(time.perf_counter() - t0, 1e-6)
    sample = int(min(nt, max(probe, probe * seconds / dt)))
    sample = max(64, sample - sample % 64)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.brute_force_simd8_parallel(T[:sample], P, softening)
        times.append(time.perf_counter() - t0)
    rate = sample * n / (sum(times) / len(times)) / 1e9
    return rate, f"first {sample} targets x all {n} sources, scaled linearly", cores, times


def cpu_barneshut_rate(P, theta, seconds):
    """Restated parallel::BarnesHut: single-thread build of the full tree + OpenMP traversal of a
    bounded target sample; the per-evaluation time is build + traversal scaled to all targets."""
    import oracle
    n = len(P)
    oracle.use_all_cores()
    cores = oracle.baseline_threads()
    t0 = time.perf_counter()
    tree = oracle.Tree(P)
    t_build = time.perf_counter() - t0
    probe = min(n, 512 * cores)
    idx = np.linspace(0, n - 1, probe).astype(np.int64)
    t0 = time.perf_counter()
    tree.traverse(P[idx, :3], theta, parallel=True)
    dt = max(time.perf_counter() - t0, 1e-6)
    sample = int(min(n, max(probe, probe * seconds / dt)))
    idx = np.linspace(0, n - 1, sample).astype(np.int64)
    t0 = time.perf_counter()
    tree.traverse(P[idx, :3], theta, parallel=True)
    t_trav = time.perf_counter() - t0
    total = t_build + t_trav * n / sample
    return (n / total, f"full single-thread build ({t_build:.1f} s) + traversal of {sample} evenly "
            f"spaced targets of {n}, scaled linearly", cores, t_build, t_trav * n / sample)


# ---- main ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bruteforce", choices=["bruteforce", "barneshut", "split"])
    ap.add_argument("--n", "--particles", dest="n", type=int, default=0,
                    help="particle count (default: BASELINE config); use --particles under torchrun, "
                         "whose own parser rejects the abbreviation --n")
    ap.add_argument("--theta", type=float, default=0.5)
    ap.add_argument("--bh-build", default="auto", choices=["auto", "replicated", "partitioned"],
                    help="multi-GPU Barnes-Hut: every GPU builds the whole tree, or one tree per GPU "
                         "over its key range joined by a top tree (PCUDA_FLAG_BH_PARTITIONED_BUILD); "
                         "auto = the library's default: partitioned from 4 GPUs on")
    ap.add_argument("--bh-route", default="auto", choices=["auto", "allgather", "alltoall"],
                    help="multi-GPU Barnes-Hut: how the accelerations reach the ranks that own the "
                         "particles (auto: all-to-all from 4 GPUs and 32M particles on)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the cpu_baseline leg and the ride-along Barnes-Hut number")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n or {"bruteforce": 1_000_000, "barneshut": 10_000_000, "split": 16_000_000}[args.workload]

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, n, world)
        return 0
    if args.workload == "bruteforce":
        return bench_bruteforce(args, n, rank, world, local_rank)
    if args.workload == "split":
        return bench_split(args, n, rank, world, local_rank)
    return bench_barneshut(args, n, rank, world, local_rank)


def reference_arm(args, n, world):
    """The reference's CPU implementation of the path (restated; kind "port") on all host cores."""
    if args.workload == "bruteforce":
        P = uniform_cloud(n)
        per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        rate, sample, cores, times = cpu_bruteforce_rate(P, per_step, steps=args.steps + args.warmup)
        times = times[args.warmup:] or times
        sample_n = int(sample.split()[1])
        value = sample_n * n / (sum(times) / len(times)) / 1e9
        line = {"impl": "reference", "metric": "brute-force pair interactions per second",
                "value": value, "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": bruteforce_config(n, world, "cpu"),
                "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                 "sample": sample + "; restated parallel::BruteForceSimd<8> "
                                 "(AVX2 rsqrt + OpenMP), oracle/baseline_simd.c"},
                "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    elif args.workload == "split":
        tgt, src = split_cloud(10_000, n)
        per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        rate, sample, cores, times = cpu_bruteforce_rate(src, per_step, steps=args.steps + args.warmup,
                                                         targets=tgt, softening=1.0)
        times = times[args.warmup:] or times
        sample_n = int(sample.split()[1])
        value = sample_n * float(len(src)) / (sum(times) / len(times)) / 1e9
        line = {"impl": "reference", "metric": "brute-force pair interactions per second",
                "value": value, "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"brute force 3-D f32 massive/massless split: 10000 massive + {n} "
                                       f"massless (BASELINE configs[2]); AccelerationSoftened::checked(1.0)",
                           "parallelism": "host threads over targets"},
                "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                 "sample": sample + "; restated parallel::BruteForceSimd<8> "
                                 "(AVX2 rsqrt + OpenMP), oracle/baseline_simd.c"},
                "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    else:
        P = plummer_cloud(n)
        value, sample, cores, tb, tt = cpu_barneshut_rate(P, args.theta, 20.0)
        line = {"impl": "reference", "metric": "Barnes-Hut particles per second (build + traversal)",
                "value": value, "unit": "particles/s", "n_gpus": world, "steps": 1, "warmup": 0,
                "ms_per_step": 1e3 * (tb + tt), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": barneshut_config(n, world, args.theta, "cpu"),
                "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores,
                                 "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def bruteforce_config(n, world, where):
    return {"workload": f"brute force 3-D f32, N={n} massive particles, all pairs "
                        f"(BASELINE configs[1]); uniform cube, mu U[1e3,1e9), seed {SEED}; "
                        f"Acceleration::checked(), softening 0",
            "n_particles": n, "pairs_per_step": n * n,
            "parallelism": f"targets sharded over {world} GPU(s), sources all-gathered (NCCL)"
            if where == "gpu" else "host threads over targets",
            "l2": "256 MiB buffer written between timed steps (L2 flush); the 16 MB source set "
                  "is re-read from L2 by design" if where == "gpu" else "n/a"}


def barneshut_config(n, world, theta, where, build="auto"):
    if build == "auto":
        build = "partitioned" if world >= 4 else "replicated"
    how = ("tree build replicated" if build == "replicated" else
           "one tree per GPU over its key range (partitioned build), trees all-gathered and joined "
           "by a top tree")
    return {"workload": f"Barnes-Hut 3-D f32 octree, theta={theta}, N={n} Plummer sphere (a=1, "
                        f"r<50a, equal mu=1/N, seed {SEED}); tree rebuilt every step "
                        f"(BASELINE configs[3]); Acceleration::checked()",
            "n_particles": n, "theta": theta,
            "parallelism": (f"{world} GPU(s): particles all-gathered (NCCL), {how}, "
                            f"targets sharded by key range" if world > 1 else "1 GPU") if where == "gpu"
            else "host threads over targets",
            "l2": "inputs + tree exceed L2 at N=10M; 256 MiB buffer written between timed steps"
            if where == "gpu" else "n/a"}


def _dist_setup(world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return dist


def _max_over_ranks(x, world, dist):
    import torch
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x, world, dist):
    import torch
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _barrier(world, dist):
    import torch
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def bench_bruteforce(args, n, rank, world, local_rank):
    import torch

    import particular_b200 as pb
    dist = _dist_setup(world, local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pb.CudaContext(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    inter = pb.Acceleration.checked()
    sh = pb.ShardedBruteForce(ctx, inter)
    P = uniform_cloud(n)
    lo, hi = pb.shard_bounds(n, world, rank)
    n_local = hi - lo
    d_local = torch.from_numpy(P[lo:hi]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)

    def timed_steps(step_fn, k):
        """Each step bracketed by events on the context stream; L2 flushed between steps."""
        times, kernel_ms, launches = [], [], 0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            ctx.sync()
            times.append(e0.elapsed_time(e1))
            t = ctx.timings()
            kernel_ms.append(t["compute_ms"])
            launches += t["kernel_launches"]
        return times, kernel_ms, launches

    # ---- device-resident value ----
    dev_step = lambda: sh.step_device(d_local, n)  # noqa: E731
    for _ in range(args.warmup):
        dev_step()
    _barrier(world, dist)
    sampler.start()
    times, kernel_ms, launches = timed_steps(dev_step, args.steps)
    _barrier(world, dist)
    sampler.stop()
    total_ms = _max_over_ranks(sum(times), world, dist)
    ms_per_step = total_ms / args.steps
    value = n * float(n) / (ms_per_step * 1e-3) / 1e9
    total_launches = int(_sum_over_ranks(launches, world, dist))

    # ---- end to end through the public API, host buffers ----
    h_local = ctx.pinned_empty((n_local, 4), np.float32)
    h_local[:] = P[lo:hi]
    h_out = ctx.pinned_empty((n_local, 3), np.float32)
    e2e_step = lambda: sh.compute_local(h_local, n, out=h_out)  # noqa: E731
    for _ in range(2):
        e2e_step()
    _barrier(world, dist)
    e2e_times, _, _ = timed_steps(e2e_step, args.steps)
    _barrier(world, dist)
    e2e_ms = _max_over_ranks(sum(e2e_times), world, dist) / args.steps
    e2e_value = n * float(n) / (e2e_ms * 1e-3) / 1e9

    # ---- roofline of the pair kernel (this rank's launch) ----
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = FLOP_PER_PAIR * n_local * float(n) / (k_ms * 1e-3) / 1e12
    sm_max_mhz = (sampler.max_mhz or ctx.sm_clock_khz / 1e3)
    peak_nominal = ctx.sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    probe_tf, _ = ctx.probe_fp32(True, 8192, 3)
    roofline = {"bound": "fp32", "kernel": "pcuda::bf::pair_kernel_f32<3,...>",
                "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
                "frac": achieved / peak_nominal,
                "peak_source": f"{ctx.sm_count} SMs x 128 FP32 lanes x 2 flop x {sm_max_mhz:.0f} MHz "
                               "(clocks.max.sm; MEASURED_PEAKS.json carries no FP32 figure)",
                "peak_probe_ffma2": probe_tf, "frac_of_probe": achieved / probe_tf,
                "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": n_local * n,
                "kernel_ms": k_ms, "traffic": None}
    if world == 1:
        roofline["traffic"], roofline["traffic_source"] = ncu_traffic("pair_kernel_f32_n1M")

    line = {"metric": "brute-force pair interactions per second", "value": value,
            "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bruteforce_config(n, world, "gpu"),
            "e2e": {"value": e2e_value, "unit": "Gpairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n_local * 16), "d2h_bytes_per_step": int(n_local * 12),
                    "bytes_are": "per rank"},
            "gpu_launches": total_launches, "roofline": roofline, "clocks": sampler.summary(),
            "device": ctx.name}

    if rank == 0 and world == 1 and not args.no_extra:
        rate, sample, cores, _ = cpu_bruteforce_rate(P, args.cpu_seconds)
        line["cpu_baseline"] = {"value": rate, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                "sample": sample + "; restated parallel::BruteForceSimd<8> (AVX2 "
                                "rsqrt + OpenMP), oracle/baseline_simd.c"}
        try:
            line["barnes_hut"] = barneshut_numbers(args, ctx, stream, flush, 10_000_000, args.theta,
                                                   steps=3, warmup=2, cpu_seconds=args.cpu_seconds)
        except Exception as e:  # keep the headline line even if the ride-along fails
            line["barnes_hut"] = {"error": repr(e)}
        try:
            line["other_configs"] = other_configs(ctx, stream, flush)
        except Exception as e:
            line["other_configs"] = {"error": repr(e)}
    if world > 1 and not args.no_extra:
        # BASELINE configs[3] at this world size rides along (collective: every rank takes part)
        try:
            bh = barneshut_numbers(args, ctx, stream, flush, 10_000_000, args.theta, steps=3, warmup=2,
                                   cpu_seconds=0.0, rank=rank, world=world, dist=dist, init_comm=False)
            line["barnes_hut"] = bh
        except Exception as e:
            line["barnes_hut"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def split_cloud(n_massive, n_massless):
    """BASELINE configs[2]: the massive bodies first, then the massless ones (the order Reordered's
    affected side has when the input is already partitioned); same draws as other_configs()."""
    rng = np.random.default_rng(SEED)
    src = uniform_cloud(n_massive)
    tgt = np.empty((n_massive + n_massless, 3), np.float32)
    tgt[:n_massive] = src[:, :3]
    tgt[n_massive:] = rng.uniform(-5e3, 5e3, (n_massless, 3))
    return tgt, src


def bench_split(args, n_massless, rank, world, local_rank):
    """BASELINE configs[2] at 1/2/4/8 GPUs: 10,000 massive act on themselves + n_massless massless
    particles; the affected particles are sharded, the massive records all-gathered each step
    (pcuda_bruteforce_f32x3_between_sharded).  Strong scaling."""
    import torch

    import particular_b200 as pb
    dist = _dist_setup(world, local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pb.CudaContext(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    n_massive = 10_000
    sb = pb.ShardedBetween(ctx, pb.AccelerationSoftened.checked(1.0))
    tgt, src = split_cloud(n_massive, n_massless)
    n_aff = len(tgt)
    lo, hi = pb.shard_bounds(n_aff, world, rank)
    slo, shi = pb.shard_bounds(n_massive, world, rank)
    d_tgt = torch.from_numpy(tgt[lo:hi]).to(dev)
    d_src = torch.from_numpy(src[slo:shi]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)

    def timed_steps(step_fn, k):
        times, launches = [], 0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            ctx.sync()
            times.append(e0.elapsed_time(e1))
            launches += ctx.timings()["kernel_launches"]
        return times, launches

    dev_step = lambda: sb.step_device(d_tgt, d_src, n_massive)  # noqa: E731
    for _ in range(args.warmup):
        dev_step()
    _barrier(world, dist)
    sampler.start()
    times, launches = timed_steps(dev_step, args.steps)
    _barrier(world, dist)
    sampler.stop()
    ms_per_step = _max_over_ranks(sum(times), world, dist) / args.steps
    pairs = float(n_aff) * n_massive
    total_launches = int(_sum_over_ranks(launches, world, dist))

    h_tgt = ctx.pinned_empty((hi - lo, 3), np.float32)
    h_tgt[:] = tgt[lo:hi]
    h_src = np.ascontiguousarray(src[slo:shi])
    h_out = ctx.pinned_empty((hi - lo, 3), np.float32)
    e2e_step = lambda: sb.compute_local(h_tgt, h_src, n_massive, out=h_out)  # noqa: E731
    for _ in range(2):
        e2e_step()
    _barrier(world, dist)
    e2e_times, _ = timed_steps(e2e_step, args.steps)
    _barrier(world, dist)
    e2e_ms = _max_over_ranks(sum(e2e_times), world, dist) / args.steps

    k_ms = sum(times) / len(times)
    achieved = FLOP_PER_PAIR * float(hi - lo) * n_massive / (k_ms * 1e-3) / 1e12
    sm_max_mhz = (sampler.max_mhz or ctx.sm_clock_khz / 1e3)
    peak_nominal = ctx.sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    line = {"metric": "brute-force pair interactions per second", "value": pairs / ms_per_step / 1e6,
            "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"brute force 3-D f32 massive/massless split: {n_massive} massive + "
                                   f"{n_massless} massless, Between(all, massive) (Reordered storage; "
                                   f"BASELINE configs[2]); uniform cube, seed {SEED}; "
                                   f"AccelerationSoftened::checked(1.0)",
                       "n_affected": n_aff, "n_affecting": n_massive, "pairs_per_step": pairs,
                       "parallelism": f"affected sharded over {world} GPU(s), massive records "
                                      f"all-gathered (NCCL, {16 * n_massive} B)",
                       "l2": "256 MiB buffer written between timed steps (L2 flush); the 160 KB source "
                             "set is re-read from L2 by design"},
            "e2e": {"value": pairs / e2e_ms / 1e6, "unit": "Gpairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h_tgt.nbytes + h_src.nbytes),
                    "d2h_bytes_per_step": int(h_out.nbytes), "bytes_are": "per rank"},
            "gpu_launches": total_launches,
            "roofline": {"bound": "fp32", "kernel": "pcuda::bf::pair_kernel_f32<3,...>",
                         "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
                         "frac": achieved / peak_nominal, "flop_per_pair": FLOP_PER_PAIR,
                         "kernel_ms": k_ms, "traffic": None,
                         "note": "whole device-resident step of this rank (fill + all-gather + pair "
                                 "kernel + split reduction)"},
            "clocks": sampler.summary(), "device": ctx.name}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def barneshut_numbers(args, ctx, stream, flush, n, theta, steps, warmup, cpu_seconds, rank=0,
                      world=1, dist=None, init_comm=True):
    """Barnes-Hut: device-resident particles/s (tree rebuilt + traversal every step), e2e through
    the host API, work counters, CPU restatement beside it.  world > 1: every rank owns a block of
    the particles, all-gathers the records, builds the identical tree and traverses its block."""
    import torch

    import particular_b200 as pb
    P = plummer_cloud(n)
    dev = torch.device("cuda", ctx.device)
    inter = pb.Acceleration.checked()
    lo, hi = pb.shard_bounds(n, world, rank)
    n_local = hi - lo
    if world == 1:
        bh = pb.BarnesHut(ctx, theta, inter)
        d_src = torch.from_numpy(P).to(dev)
        d_out = torch.empty((n, 3), dtype=torch.float32, device=dev)
        dev_step = lambda: bh.compute_device(None, n, d_src.data_ptr(), n, d_out.data_ptr())  # noqa: E731
    else:
        bh = pb.ShardedBarnesHut(ctx, theta, inter, init_comm=init_comm)  # False: the context's
        bh.world, bh.rank = world, rank                                   # communicator exists already
        d_src = torch.from_numpy(P[lo:hi]).to(dev)
        dev_step = lambda: bh.step_device(d_src, n)  # noqa: E731

    def run(step_fn, k):
        times, tm, launches = [], [], 0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step_fn()
            e1.record(stream)
            ctx.sync()
            times.append(e0.elapsed_time(e1))
            t = ctx.timings()
            tm.append(t)
            launches += t["kernel_launches"]
        return times, tm, launches

    for _ in range(warmup):
        dev_step()
    ctx.sync()
    if dist is not None:
        _barrier(world, dist)
    times, tm, launches = run(dev_step, steps)
    if dist is not None:
        _barrier(world, dist)
    ms = _max_over_ranks(sum(times), world, dist) / steps if dist is not None else sum(times) / steps
    build_ms = sum(t["build_ms"] for t in tm) / len(tm)
    trav_ms = sum(t["compute_ms"] for t in tm) / len(tm)
    comm_ms = sum(t["comm_ms"] for t in tm) / len(tm)
    plain = pb.BarnesHut(ctx, theta, inter)
    counters = plain.last_counters()
    h_in = ctx.pinned_empty((n_local, 4), np.float32)
    h_in[:] = P[lo:hi]
    h_out = ctx.pinned_empty((n_local, 3), np.float32)
    if world == 1:
        e2e_step = lambda: bh.compute(h_in, out=h_out)  # noqa: E731
    else:
        e2e_step = lambda: bh.compute_local(h_in, n, out=h_out)  # noqa: E731
    e2e_step()
    if dist is not None:
        _barrier(world, dist)
    e2e_times, _, _ = run(e2e_step, steps)
    if dist is not None:
        _barrier(world, dist)
    e2e_ms = (_max_over_ranks(sum(e2e_times), world, dist) if dist is not None else sum(e2e_times)) / steps
    inter_n = counters["node_interactions"] + counters["particle_interactions"]
    # algorithmic bytes of the traversal: one 32-byte record per node test, one
    # 16-byte record per particle entry appended to a group's list, 16 B read + 12 B written per target
    groups = max(counters.get("groups", 0), 1)
    part_entries = counters["particle_interactions"] * groups / max(n_local, 1)
    trav_bytes = 32.0 * counters["node_tests"] + 16.0 * part_entries + 28.0 * n_local
    hbm_peak, hbm_src = measured_hbm_peak()
    achieved_gbs = trav_bytes / (trav_ms * 1e-3) / 1e9
    total_launches = int(_sum_over_ranks(launches, world, dist)) if dist is not None else launches
    out = {"metric": "Barnes-Hut particles per second (build + traversal)",
           "value": n / (ms * 1e-3), "unit": "particles/s", "ms_per_step": ms,
           "comm_ms": comm_ms, "build_ms": build_ms, "traverse_ms": trav_ms, "steps": steps,
           "warmup": warmup, "config": barneshut_config(n, world, theta, "gpu", args.bh_build),
           "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "particles/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": n_local * 16, "d2h_bytes_per_step": n_local * 12,
                   "bytes_are": "per rank"},
           "gpu_launches": total_launches, "counters_last_step_rank0": counters,
           "roofline": {"bound": "hbm", "kernel": "pcuda::bh::traverse2_kernel", "achieved": achieved_gbs,
                        "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                        "peak_source": hbm_src, "bytes_per_launch": trav_bytes, "kernel_ms": trav_ms,
                        "traffic": ncu_traffic("traverse2_kernel_n10M")[0] if world == 1 and n == 10_000_000 else None,
                        "note": "the node set is served from L2 (ncu: 93 % hit rate, ~1.1 GB of DRAM "
                                "traffic per launch); the kernel is FP32-pipe / issue bound, see "
                                "traversal_fp32 and profiles/"},
           "traversal_fp32": {"achieved": FLOP_PER_PAIR * inter_n / (trav_ms * 1e-3) / 1e12,
                              "peak": ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12,
                              "frac": FLOP_PER_PAIR * inter_n / (trav_ms * 1e-3) / 1e12
                              / (ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12),
                              "unit": "TFLOP/s", "interactions_per_target": inter_n / max(n_local, 1),
                              "note": "20 flop per accepted interaction, this rank's targets; the binding "
                                      "resource (DESIGN.md K5): issue cycles, 2 per packed FP32 instruction"}}
    if not args.no_extra and rank == 0 and world == 1:
        rate, sample, cores, tb, tt = cpu_barneshut_rate(P, theta, cpu_seconds)
        out["cpu_baseline"] = {"value": rate, "unit": "particles/s", "cores": cores, "kind": "port",
                               "sample": sample + "; restated parallel::BarnesHut",
                               "build_s": tb, "traverse_s": tt}
    del d_src
    return out


def bench_barneshut(args, n, rank, world, local_rank):
    import torch

    import particular_b200 as pb
    dist = _dist_setup(world, local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pb.CudaContext(local_rank, partitioned_build={"auto": None, "partitioned": True,
                                                        "replicated": False}[args.bh_build])
    from particular_b200._ffi import lib as _lib
    assert _lib.pcuda_debug_set(b"bh_route", {"auto": 0, "allgather": 1, "alltoall": 2}[args.bh_route]) == 0
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = barneshut_numbers(args, ctx, stream, flush, n, args.theta, args.steps, args.warmup,
                            args.cpu_seconds, rank, world, dist)
    sampler.stop()
    line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": res["config"], "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"], "clocks": sampler.summary(), "device": ctx.name,
            "comm_ms": res["comm_ms"], "build_ms": res["build_ms"], "traverse_ms": res["traverse_ms"],
            "counters_last_step_rank0": res["counters_last_step_rank0"],
            "roofline": res["roofline"], "traversal_fp32": res["traversal_fp32"]}
    if "cpu_baseline" in res:
        line["cpu_baseline"] = res["cpu_baseline"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
