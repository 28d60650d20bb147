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



# ---- parity of the measured result buffers (checker only: the oracle is never on the timed path) ---
PARITY_ROWS = 1024


def _rel(a, ref):
    den = np.linalg.norm(ref, axis=1)
    return np.linalg.norm(np.asarray(a, np.float64) - ref, axis=1) / np.where(den > 0, den, 1.0)


def _sample_rows(n_rows, m=PARITY_ROWS):
    return np.unique(np.linspace(0, n_rows - 1, min(m, n_rows)).astype(np.int64))


def _oracle_threads(world):
    import oracle
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    oracle.lib().oracle_set_threads(max(1, cores // max(world, 1)))


def parity_bruteforce(targets, rows, got, sources, softening, world=1):
    """`rows` of the measured buffer `got` against the extended-precision sum over all `sources`
    (oracle.brute_force_exact) and, beside it, the error of the restated sequential::BruteForce f32
    left fold itself (err_gpu <= max(1e-5, err_ref) at N = 1M)."""
    import oracle
    _oracle_threads(world)
    t = np.ascontiguousarray(targets[rows])
    exact = oracle.brute_force_exact(t, sources, softening)
    ref = oracle.brute_force_parallel(t, sources, softening)
    e_gpu, e_ref = _rel(got[rows], exact), _rel(ref, exact)
    return e_gpu, e_ref


def parity_summary(e_gpu, e_ref, kind):
    e_gpu, e_ref = np.asarray(e_gpu), np.asarray(e_ref)
    out = {"n": int(len(e_gpu)), "max_rel": float(e_gpu.max()), "p99": float(np.percentile(e_gpu, 99)),
           "median": float(np.median(e_gpu)), "ref_max_rel": float(e_ref.max()),
           "ref_p99": float(np.percentile(e_ref, 99)), "ref_median": float(np.median(e_ref))}
    if kind == "bruteforce":
        out["against"] = ("extended-precision sum (oracle.brute_force_exact); ref_* = the restated "
                          "sequential::BruteForce f32 fold on the same rows")
        out["ok"] = bool(e_gpu.max() <= max(1e-5, e_ref.max()))
    else:
        out["against"] = ("extended-precision sum; ref_* = the restated sequential::BarnesHut at the "
                          "same theta on the same rows; bound 1.1 x ref + 2e-6 on median / p99 / max")
        out["ok"] = bool(out["median"] <= 1.1 * out["ref_median"] + 2e-6 and
                         out["p99"] <= 1.1 * out["ref_p99"] + 2e-6 and
                         out["max_rel"] <= 1.1 * out["ref_max_rel"] + 2e-6)
    return out


def gather_errors(e, world, dist):
    """Every rank's sampled errors on rank 0 (equal-length arrays)."""
    import torch
    if world == 1:
        return e
    t = torch.from_numpy(np.ascontiguousarray(e, np.float64)).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.cat(parts).cpu().numpy()


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
    # configs[4a]: f64 precision path
    n64 = 262_144
    p64 = uniform_cloud(n64).astype(np.float64)
    d_p = torch.from_numpy(p64).to(dev)
    d_o = torch.empty((n64, 3), dtype=torch.float64, device=dev)
    bf64 = pb.BruteForce(ctx, pb.Acceleration.checked())
    ms = timed(lambda: bf64.compute_device(None, n64, d_p.data_ptr(), n64, d_o.data_ptr(), "f64x3"))
    peak64 = ctx.sm_count * 64 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    tf = FLOP_PER_PAIR * float(n64) * n64 / (ms * 1e-3) / 1e12
    rows = _sample_rows(n64)
    got64 = d_o.cpu().numpy()
    import oracle
    _oracle_threads(1)
    ref64 = oracle.brute_force_parallel(np.ascontiguousarray(p64[rows, :3]), p64)
    e64 = _rel(got64[rows], ref64)
    out["f64_256k"] = {"config": "BASELINE configs[4]: brute force 3-D f64, N=262144, Acceleration::checked()",
                       "ms_per_step": ms, "value": float(n64) * n64 / ms / 1e6, "unit": "Gpairs/s",
                       "fp64_tflops_20flop_per_pair": tf, "fp64_peak_tflops_nominal": peak64,
                       "frac": tf / peak64,
                       "parity": {"n": int(len(rows)), "max_rel": float(e64.max()), "bound": 1e-12,
                                  "ok": bool(e64.max() <= 1e-12),
                                  "against": "restated sequential::BruteForce f64 fold, same rows"}}
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
    rows = _sample_rows(n2, 512)
    got2 = d_o.cpu().numpy()
    ex2 = oracle.brute_force_exact(np.ascontiguousarray(p2[rows, :2]), p2, 100.0)
    e2 = _rel(got2[rows], ex2)
    out["barnes_hut_2d_4M"] = {"config": "BASELINE configs[4]: Barnes-Hut 2-D f32 quadtree, theta=0.5, N=4194304 "
                                         "uniform square, AccelerationSoftened::checked(100)",
                               "ms_per_step": ms, "build_ms": t["build_ms"], "traverse_ms": t["compute_ms"],
                               "value": n2 / (ms * 1e-3), "unit": "particles/s",
                               "theta_error_vs_exact": {"n": int(len(rows)), "median": float(np.median(e2)),
                                                        "p99": float(np.percentile(e2, 99)),
                                                        "max_rel": float(e2.max())}}
    del d_p, d_o

    # the gravity pair term written as a USER-DEFINED interaction (NVRTC, IEEE sqrt and
    # division, no fused multiply-add: bit-identical to the CPU fold) next to the hand-written kernel
    try:
        nc = 131_072
        pc = uniform_cloud(nc)
        custom = pb.CustomInteraction(
            ctx, CUSTOM_GRAVITY_SRC, np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4")]),
            np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("mu", "f4")]),
            np.dtype([("ax", "f4"), ("ay", "f4"), ("az", "f4")]), np.dtype([("softening", "f4")]), push=(0.0,))
        aff = np.ascontiguousarray(pc[:, :3]).view(custom.affected_dtype).reshape(-1)
        src = pc.view(custom.affecting_dtype).reshape(-1)
        custom.brute_force(aff, src)
        t0 = time.perf_counter()
        got = custom.brute_force(aff, src)
        wall = time.perf_counter() - t0
        k_ms = ctx.timings()["compute_ms"]
        rows = _sample_rows(nc, 256)
        refc = oracle.brute_force(pc[rows, :3], pc)
        gotc = np.stack([got["ax"], got["ay"], got["az"]], axis=1)
        bf32 = pb.BruteForce(ctx, pb.Acceleration.checked())
        d_pc = torch.from_numpy(pc).to(dev)
        d_oc = torch.empty((nc, 3), dtype=torch.float32, device=dev)
        k1_ms = timed(lambda: bf32.compute_device(None, nc, d_pc.data_ptr(), nc, d_oc.data_ptr()))
        out["custom_gravity_128k"] = {
            "config": "pcuda_interaction_* (CustomInteraction): Acceleration::checked written as user source, "
                      "3-D f32, N=131072; kernel = TMA ring + 2 affected per thread (csrc/custom.cu)",
            "kernel_ms": k_ms, "value": float(nc) * nc / k_ms / 1e6, "unit": "Gpairs/s",
            "host_call_ms": 1e3 * wall, "hand_written_kernel_ms": k1_ms,
            "hand_written_value": float(nc) * nc / k1_ms / 1e6,
            "bit_identical_to_cpu_fold": bool(np.array_equal(gotc[rows], refc))}
        custom.close()
        del d_pc, d_oc
    except Exception as e:
        out["custom_gravity_128k"] = {"error": repr(e)}

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
    res["cpu_scalar"] = cpu_scalar_1024()
    got = one.compute(pb_small)
    res["parity"] = {"n": nb, "max_rel": float(_rel(got, oracle.brute_force(pb_small[:, :3], pb_small)
                                                   .astype(np.float64)).max()),
                     "bound": 1e-5, "against": "restated sequential::BruteForce (bit-faithful f32 fold)"}
    res["parity"]["ok"] = bool(res["parity"]["max_rel"] <= 1e-5)
    out["stepping_1024"] = {"config": "BASELINE configs[0] shape (benches/benchmark.rs:97-131, N=1024): "
                                      "device-resident stepping (pcuda_sim_*), brute force 3-D f32, 2048 steps; "
                                      "one_shot_host_api = upload + kernel + read-back per step, as the "
                                      "reference's wgpu operator works; cpu_scalar = the restated "
                                      "sequential::BruteForce on one host core",
                            **res}
    return out


CUSTOM_GRAVITY_SRC = """
struct Affected { float x, y, z; };
struct Affecting { float x, y, z, mu; };
struct Interaction { float ax, ay, az; };
struct Push { float softening; };
__device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out) {
    const float dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
    const float n = dx * dx + dy * dy + dz * dz;
    if (n != 0.f) {
        const float ns = n + push.softening * push.softening;
        const float s = p2.mu / (ns * sqrtf(ns));
        out.ax += dx * s; out.ay += dy * s; out.az += dz * s;
    }
}
"""


# ---- CPU arms ----------------------------------------------------------------------------------------
def cpu_scalar_1024(n=1024, reps=20):
    """BASELINE configs[0]: sequential::BruteForce (scalar, one core) at the criterion bench shape."""
    import oracle
    P = uniform_cloud(n)
    oracle.brute_force(P[:, :3], P)
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.brute_force(P[:, :3], P)
    dt = (time.perf_counter() - t0) / reps
    return {"ms_per_step": 1e3 * dt, "value": n * n / dt / 1e9, "unit": "Gpairs/s", "cores": 1, "n": n,
            "kind": "port", "what": "restated sequential::BruteForce, oracle/oracle_impl.inc"}


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


def cpu_barneshut_rate(P, theta, seconds, tree=None, t_build=None):
    """Restated parallel::BarnesHut: single-thread build of the full tree + OpenMP traversal of a
    bounded target sample; the per-evaluation time is build + traversal scaled to all targets.
    `tree` / `t_build`: a tree over P built (and timed) by the caller already."""
    import oracle
    n = len(P)
    oracle.use_all_cores()
    cores = oracle.baseline_threads()
    if tree is None:
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
    ap.add_argument("--bh-build", default="auto", choices=["auto", "replicated", "partitioned", "let"],
                    help="multi-GPU Barnes-Hut: every GPU builds the whole tree; one tree per GPU over its "
                         "key range, all trees all-gathered (PCUDA_FLAG_BH_PARTITIONED_BUILD); or locally "
                         "essential trees (PCUDA_FLAG_BH_LET_BUILD) — auto = the library's default: LET")
    ap.add_argument("--bh-route", default="auto", choices=["auto", "allgather", "alltoall"],
                    help="multi-GPU Barnes-Hut: how the accelerations reach the ranks that own the "
                         "particles (auto: all-to-all from 4 GPUs and 32M particles on)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the cpu_baseline leg and the ride-along Barnes-Hut / split numbers")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the result buffers")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if os.environ.get("PCUDA_DEBUG") and args.impl == "ours":  # tuning hooks, e.g. PCUDA_DEBUG=bh_let_trace=1
        from particular_b200._ffi import lib as _dbg
        for kv in os.environ["PCUDA_DEBUG"].split(","):
            k, v = kv.split("=")
            assert _dbg.pcuda_debug_set(k.encode(), int(v)) == 0, kv
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


SIMD_NOTE = "; restated parallel::BruteForceSimd<8> (AVX2 rsqrt + OpenMP), oracle/baseline_simd.c"


def reference_arm(args, n, world):
    """The reference's CPU implementation of the path (restated; kind "port") on all host cores.
    The brute-force line also carries the other reference numbers a ratio can be formed from:
    `scalar_1024` (BASELINE configs[0], sequential::BruteForce at N = 1024, one core) and
    `barnes_hut` (restated parallel::BarnesHut, configs[3]) — bounded samples, stated."""
    common = {"impl": "reference", "n_gpus": world, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": 0}
    if args.workload in ("bruteforce", "split"):
        if args.workload == "bruteforce":
            P, tgt, soft = uniform_cloud(n), None, 0.0
            n_src, config = n, bruteforce_config(n, world)
        else:
            tgt, P = split_cloud(10_000, n)
            soft, n_src, config = 1.0, len(P), split_config(10_000, n, world)
        per_step = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        rate, sample, cores, times = cpu_bruteforce_rate(P, per_step, steps=args.steps + args.warmup,
                                                         targets=tgt, softening=soft)
        times = times[args.warmup:] or times
        sample_n = int(sample.split()[1])
        value = sample_n * float(n_src) / (sum(times) / len(times)) / 1e9
        line = {**common, "metric": "brute-force pair interactions per second", "value": value,
                "unit": "Gpairs/s", "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(times) / len(times), "config": config,
                "ms_per_step_is": "the time of the stated target sample, not of a whole step",
                "cpu_baseline": {"value": value, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                 "sample": sample + SIMD_NOTE},
                "e2e": {"value": value, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if args.workload == "bruteforce" and not args.no_extra:
            line["scalar_1024"] = cpu_scalar_1024()
            try:
                v, smp, c, tb, tt = cpu_barneshut_rate(plummer_cloud(10_000_000), args.theta, 8.0)
                line["barnes_hut"] = {"value": v, "unit": "particles/s", "cores": c, "kind": "port",
                                      "build_s": tb, "traverse_s": tt, "sample": smp}
            except Exception as e:
                line["barnes_hut"] = {"error": repr(e)}
    else:
        P = plummer_cloud(n)
        value, sample, cores, tb, tt = cpu_barneshut_rate(P, args.theta, 20.0)
        line = {**common, "metric": "Barnes-Hut particles per second (build + traversal)", "value": value,
                "unit": "particles/s", "steps": 1, "warmup": 0, "ms_per_step": 1e3 * (tb + tt),
                "config": barneshut_config(n, args.theta, world, args.bh_build),
                "cpu_baseline": {"value": value, "unit": "particles/s", "cores": cores, "kind": "port",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "particles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# `config` is the same object in both arms (the driver compares them): it names the workload and says,
# for either arm, how it is spread and what happens to the L2 between timed steps.
def bruteforce_config(n, world):
    return {"workload": f"brute force 3-D f32, N={n} massive particles, all pairs "
                        f"(BASELINE configs[1]); uniform cube, mu U[1e3,1e9), seed {SEED}; "
                        f"Acceleration::checked(), softening 0",
            "n_particles": n, "pairs_per_step": n * n,
            "parallelism": f"GPU arm: targets sharded over {world} GPU(s), sources all-gathered (NCCL); "
                           f"reference arm: host threads over targets",
            "l2": "GPU arm: 256 MiB buffer written between timed steps (L2 flush), the 16 MB source set is "
                  "re-read from L2 by design; reference arm: n/a"}


def split_config(n_massive, n_massless, world):
    return {"workload": f"brute force 3-D f32 massive/massless split: {n_massive} massive + {n_massless} "
                        f"massless, Between(all, massive) (Reordered storage; BASELINE configs[2]); "
                        f"uniform cube, seed {SEED}; AccelerationSoftened::checked(1.0)",
            "n_affected": n_massive + n_massless, "n_affecting": n_massive,
            "pairs_per_step": float(n_massive + n_massless) * n_massive,
            "parallelism": f"GPU arm: affected sharded over {world} GPU(s), massive records all-gathered "
                           f"(NCCL, {16 * n_massive} B); reference arm: host threads over targets",
            "l2": "GPU arm: 256 MiB buffer written between timed steps (L2 flush), the 160 KB source set is "
                  "re-read from L2 by design; reference arm: n/a"}


def barneshut_config(n, theta, world, build="auto"):
    if build == "auto":
        build = "let" if world >= 3 else "replicated"
    how = {"replicated": "particles all-gathered, tree build replicated",
           "partitioned": "particles all-gathered, one tree per GPU over its key range (partitioned build), "
                          "trees all-gathered and joined by a top tree",
           "let": "particles sent to the owners of their key ranges, one tree per GPU, locally essential trees "
                  "exchanged (all-to-all) and joined by a top tree"}[build]
    return {"workload": f"Barnes-Hut 3-D f32 octree, theta={theta}, N={n} Plummer sphere (a=1, "
                        f"r<50a, equal mu=1/N, seed {SEED}); tree rebuilt every step "
                        f"(BASELINE configs[3]); Acceleration::checked()",
            "n_particles": n, "theta": theta,
            "parallelism": (f"GPU arm: {world} GPU(s), {how}, targets sharded by key range" if world > 1
                            else "GPU arm: 1 GPU") + "; reference arm: single-thread build, host threads "
                                                     "over targets",
            "l2": "GPU arm: inputs + tree exceed L2 at N=10M, 256 MiB buffer written between timed steps; "
                  "reference arm: n/a"}


def _dist_setup(world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return dist


def _max_over_ranks(x, world, dist):
    import torch
    if world == 1 or dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x, world, dist):
    import torch
    if world == 1 or dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _barrier(world, dist):
    import torch
    torch.cuda.synchronize()
    if world > 1 and dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def _timed_steps(ctx, stream, flush, step_fn, k):
    """k steps, each bracketed by CUDA events on the context stream; L2 flushed between steps."""
    import torch
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


def _compact(d, keys):
    return {k: (float(f"{d[k]:.3g}") if isinstance(d[k], float) else d[k]) for k in keys if k in d}


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

    # ---- device-resident value ----
    dev_step = lambda: sh.step_device(d_local, n)  # noqa: E731
    for _ in range(args.warmup):
        dev_step()
    _barrier(world, dist)
    sampler.start()
    times, tm, launches = _timed_steps(ctx, stream, flush, dev_step, args.steps)
    _barrier(world, dist)
    sampler.stop()
    kernel_ms = [t["compute_ms"] for t in tm]
    total_ms = _max_over_ranks(sum(times), world, dist)
    ms_per_step = total_ms / args.steps
    value = n * float(n) / (ms_per_step * 1e-3) / 1e9
    total_launches = int(_sum_over_ranks(launches, world, dist))
    spread = {"ms_min": min(times), "ms_median": sorted(times)[len(times) // 2], "ms_max": max(times),
              "of": "rank 0's timed steps (criterion-style spread; `ms_per_step` is the max-over-ranks mean)"}

    # ---- end to end through the public API, host buffers ----
    h_local = ctx.pinned_empty((n_local, 4), np.float32)
    h_local[:] = P[lo:hi]
    h_out = ctx.pinned_empty((n_local, 3), np.float32)
    e2e_step = lambda: sh.compute_local(h_local, n, out=h_out)  # noqa: E731
    for _ in range(2):
        e2e_step()
    _barrier(world, dist)
    e2e_times, _, _ = _timed_steps(ctx, stream, flush, e2e_step, args.steps)
    _barrier(world, dist)
    e2e_ms = _max_over_ranks(sum(e2e_times), world, dist) / args.steps
    e2e_value = n * float(n) / (e2e_ms * 1e-3) / 1e9

    # ---- parity of the result buffer the e2e steps wrote (every rank: rows of its own block) ----
    parity = None
    if not args.no_parity:
        rows = _sample_rows(n_local, max(128, PARITY_ROWS // world))
        e_gpu, e_ref = parity_bruteforce(P[lo:hi, :3], rows, np.asarray(h_out), P, 0.0, world)
        parity = parity_summary(gather_errors(e_gpu, world, dist), gather_errors(e_ref, world, dist),
                                "bruteforce")
        parity["rows"] = f"{len(rows)} evenly spaced rows of every rank's block, e2e result buffer"

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
            "config": bruteforce_config(n, world),
            "e2e": {"value": e2e_value, "unit": "Gpairs/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(n_local * 16), "d2h_bytes_per_step": int(n_local * 12),
                    "bytes_are": "per rank"},
            "gpu_launches": total_launches, "roofline": roofline, "clocks": sampler.summary(),
            "device": ctx.name, "step_spread": spread}

    if rank == 0 and world == 1 and not args.no_extra:
        rate, sample, cores, _ = cpu_bruteforce_rate(P, args.cpu_seconds)
        line["cpu_baseline"] = {"value": rate, "unit": "Gpairs/s", "cores": cores, "kind": "port",
                                "sample": sample + SIMD_NOTE}
    del d_local
    bh = split = None
    if not args.no_extra:
        # BASELINE configs[3] and configs[2] ride along at every world size (collective: every rank takes part)
        try:
            bh = barneshut_numbers(args, ctx, stream, flush, 10_000_000, args.theta, steps=3, warmup=2,
                                   cpu_seconds=args.cpu_seconds, rank=rank, world=world, dist=dist,
                                   init_comm=False)
        except Exception as e:  # keep the headline line even if a ride-along fails
            bh = {"error": repr(e)}
        try:
            split = split_numbers(args, ctx, stream, flush, 10_000, 16_000_000, steps=3, warmup=2,
                                  rank=rank, world=world, dist=dist, init_comm=False)
        except Exception as e:
            split = {"error": repr(e)}
        line["barnes_hut"], line["split_10k_massive_16M_massless"] = bh, split
        if rank == 0 and world == 1:
            try:
                line["other_configs"] = other_configs(ctx, stream, flush)
            except Exception as e:
                line["other_configs"] = {"error": repr(e)}
    # compact copies: last in the line (the driver keeps the tail of stdout) and inside `roofline`
    # (a key the driver's parser keeps whole)
    ride = {}
    if parity is not None:
        ride["parity"] = _compact(parity, ["n", "max_rel", "p99", "ref_max_rel", "ok"])
    if bh is not None:
        ride["bh"] = bh.get("compact", bh)
    if split is not None:
        ride["split"] = split.get("compact", split)
    roofline["ride_along"] = ride
    if parity is not None:
        line["parity"] = parity
    for k, v in ride.items():
        if k != "parity":
            line[k] = v
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def split_cloud(n_massive, n_massless):
    """BASELINE configs[2]: the massive bodies first, then the massless ones (the order Reordered's
    affected side has when the input is already partitioned)."""
    rng = np.random.default_rng(SEED)
    src = uniform_cloud(n_massive)
    tgt = np.empty((n_massive + n_massless, 3), np.float32)
    tgt[:n_massive] = src[:, :3]
    tgt[n_massive:] = rng.uniform(-5e3, 5e3, (n_massless, 3))
    return tgt, src


def split_numbers(args, ctx, stream, flush, n_massive, n_massless, steps, warmup, rank=0, world=1,
                  dist=None, init_comm=True):
    """BASELINE configs[2] at 1/2/4/8 GPUs: n_massive massive bodies act on themselves + n_massless
    massless particles; the affected particles are sharded, the massive records all-gathered each
    step (pcuda_bruteforce_f32x3_between_sharded).  Strong scaling."""
    import torch

    import particular_b200 as pb
    dev = torch.device("cuda", ctx.device)
    sb = pb.ShardedBetween(ctx, pb.AccelerationSoftened.checked(1.0), init_comm=init_comm and world > 1)
    sb.world, sb.rank = world, rank
    tgt, src = split_cloud(n_massive, n_massless)
    n_aff = len(tgt)
    lo, hi = pb.shard_bounds(n_aff, world, rank)
    slo, shi = pb.shard_bounds(n_massive, world, rank)
    d_tgt = torch.from_numpy(tgt[lo:hi]).to(dev)
    d_src = torch.from_numpy(src[slo:shi]).to(dev)
    dev_step = lambda: sb.step_device(d_tgt, d_src, n_massive)  # noqa: E731
    for _ in range(warmup):
        dev_step()
    _barrier(world, dist)
    times, tm, launches = _timed_steps(ctx, stream, flush, dev_step, steps)
    _barrier(world, dist)
    ms_per_step = _max_over_ranks(sum(times), world, dist) / steps
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
    e2e_times, _, _ = _timed_steps(ctx, stream, flush, e2e_step, steps)
    _barrier(world, dist)
    e2e_ms = _max_over_ranks(sum(e2e_times), world, dist) / steps

    parity = None
    if not args.no_parity:
        rows = _sample_rows(hi - lo, max(128, PARITY_ROWS // world))
        e_gpu, e_ref = parity_bruteforce(tgt[lo:hi], rows, np.asarray(h_out), src, 1.0, world)
        parity = parity_summary(gather_errors(e_gpu, world, dist), gather_errors(e_ref, world, dist),
                                "bruteforce")
    k_ms = sum(t["compute_ms"] for t in tm) / len(tm)
    achieved = FLOP_PER_PAIR * float(hi - lo) * n_massive / (k_ms * 1e-3) / 1e12
    peak_nominal = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    out = {"metric": "brute-force pair interactions per second", "value": pairs / ms_per_step / 1e6,
           "unit": "Gpairs/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "config": split_config(n_massive, n_massless, world),
           "e2e": {"value": pairs / e2e_ms / 1e6, "unit": "Gpairs/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": int(h_tgt.nbytes + h_src.nbytes),
                   "d2h_bytes_per_step": int(h_out.nbytes), "bytes_are": "per rank"},
           "gpu_launches": total_launches,
           "roofline": {"bound": "fp32", "kernel": "pcuda::bf::pair_kernel_f32<3,...>",
                        "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
                        "frac": achieved / peak_nominal, "flop_per_pair": FLOP_PER_PAIR,
                        "kernel_ms": k_ms, "traffic": None}}
    if parity is not None:
        out["parity"] = parity
    out["compact"] = {"value": round(out["value"], 1), "unit": "Gpairs/s", "ms": round(ms_per_step, 3),
                      "e2e": round(out["e2e"]["value"], 1), "e2e_ms": round(e2e_ms, 3),
                      "frac": round(achieved / peak_nominal, 4)}
    if parity is not None:
        out["compact"]["parity"] = _compact(parity, ["n", "max_rel", "ref_max_rel", "ok"])
    del d_tgt, d_src
    return out


def bench_split(args, n_massless, rank, world, local_rank):
    import torch

    import particular_b200 as pb
    dist = _dist_setup(world, local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pb.CudaContext(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = split_numbers(args, ctx, stream, flush, 10_000, n_massless, args.steps, args.warmup, rank, world, dist)
    sampler.stop()
    line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": res["config"], "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "clocks": sampler.summary(),
            "device": ctx.name}
    if "parity" in res:
        line["parity"] = res["parity"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def barneshut_numbers(args, ctx, stream, flush, n, theta, steps, warmup, cpu_seconds, rank=0,
                      world=1, dist=None, init_comm=True):
    """Barnes-Hut: device-resident particles/s (tree rebuilt + traversal every step), e2e through
    the host API, work counters, parity of the e2e result buffer, CPU restatement beside it.
    world > 1: every rank owns a block of the particles (pcuda_barneshut_f32x3_sharded)."""
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

    for _ in range(warmup):
        dev_step()
    ctx.sync()
    _barrier(world, dist)
    times, tm, launches = _timed_steps(ctx, stream, flush, dev_step, steps)
    _barrier(world, dist)
    ms = _max_over_ranks(sum(times), world, dist) / steps
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
    _barrier(world, dist)
    e2e_times, _, _ = _timed_steps(ctx, stream, flush, e2e_step, steps)
    _barrier(world, dist)
    e2e_ms = _max_over_ranks(sum(e2e_times), world, dist) / steps
    inter_n = counters["node_interactions"] + counters["particle_interactions"]
    # algorithmic bytes of the traversal: one 32-byte record per node test, one
    # 16-byte record per particle entry appended to a group's list, 16 B read + 12 B written per target
    groups = max(counters.get("groups", 0), 1)
    part_entries = counters["particle_interactions"] * groups / max(n_local, 1)
    trav_bytes = 32.0 * counters["node_tests"] + 16.0 * part_entries + 28.0 * n_local
    hbm_peak, hbm_src = measured_hbm_peak()
    achieved_gbs = trav_bytes / (trav_ms * 1e-3) / 1e9
    total_launches = int(_sum_over_ranks(launches, world, dist))
    fp32_peak = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
    fp32_ach = FLOP_PER_PAIR * inter_n / (trav_ms * 1e-3) / 1e12

    # ---- parity of the e2e result buffer: every rank against the extended-precision sum; rank 0 also
    # runs the restated reference algorithm (sequential::BarnesHut, same theta) on its rows ----
    parity, cpu_base = None, None
    want_cpu = not args.no_extra and rank == 0 and world == 1 and cpu_seconds > 0
    if not args.no_parity:
        import oracle
        _oracle_threads(world)
        rows = _sample_rows(n_local, max(128, PARITY_ROWS // world))
        t_rows = np.ascontiguousarray(P[lo:hi][rows, :3])
        exact = oracle.brute_force_exact(t_rows, P)
        e_gpu = gather_errors(_rel(np.asarray(h_out)[rows], exact), world, dist)
        if rank == 0:
            oracle.use_all_cores()
            t0 = time.perf_counter()
            tree = oracle.Tree(P)
            t_build = time.perf_counter() - t0
            e_ref = _rel(tree.traverse(t_rows, theta, parallel=True), exact)
            parity = parity_summary(e_gpu, e_ref, "barneshut")
            parity["rows"] = (f"{len(rows)} evenly spaced rows of every rank's block (e2e result buffer); "
                              f"reference algorithm on rank 0's {len(rows)} rows")
            if want_cpu:
                cpu_base = cpu_barneshut_rate(P, theta, cpu_seconds, tree=tree, t_build=t_build)
            del tree
        _barrier(world, dist)
    elif want_cpu:
        cpu_base = cpu_barneshut_rate(P, theta, cpu_seconds)

    out = {"metric": "Barnes-Hut particles per second (build + traversal)",
           "value": n / (ms * 1e-3), "unit": "particles/s", "ms_per_step": ms,
           "comm_ms": comm_ms, "build_ms": build_ms, "traverse_ms": trav_ms, "steps": steps,
           "warmup": warmup, "config": barneshut_config(n, theta, world, args.bh_build),
           "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "particles/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": n_local * 16, "d2h_bytes_per_step": n_local * 12,
                   "bytes_are": "per rank"},
           "gpu_launches": total_launches, "counters_last_step_rank0": counters,
           "roofline": {"bound": "fp32/issue", "kernel": "pcuda::bh::traverse2_kernel",
                        "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak,
                        "peak_source": f"{ctx.sm_count} SMs x 128 FP32 lanes x 2 flop x "
                                       f"{ctx.sm_clock_khz / 1e3:.0f} MHz (nominal)",
                        "flop_per_interaction": FLOP_PER_PAIR,
                        "interactions_per_target": inter_n / max(n_local, 1), "kernel_ms": trav_ms,
                        "traffic": ncu_traffic("traverse2_kernel_n10M")[0] if world == 1 and n == 10_000_000 else None,
                        "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                                "bytes_per_launch": trav_bytes,
                                "note": "algorithmic node / particle record bytes; the node set is served from "
                                        "L2 (ncu: ~93 % hit rate, ~1.1 GB of DRAM traffic per launch), so HBM is "
                                        "not the binding roof"},
                        "note": "20 flop per accepted interaction, this rank's targets; the binding resource "
                                "(DESIGN.md K5) is issue cycles: 2 per packed FP32 instruction"}}
    if parity is not None:
        out["parity"] = parity
    if cpu_base is not None:
        rate, sample, cores, tb, tt = cpu_base
        out["cpu_baseline"] = {"value": rate, "unit": "particles/s", "cores": cores, "kind": "port",
                               "sample": sample + "; restated parallel::BarnesHut",
                               "build_s": tb, "traverse_s": tt}
    out["compact"] = {"value": round(out["value"]), "unit": "particles/s", "ms": round(ms, 3),
                      "comm": round(comm_ms, 3), "build": round(build_ms, 3), "trav": round(trav_ms, 3),
                      "e2e": round(out["e2e"]["value"]), "e2e_ms": round(e2e_ms, 3)}
    if parity is not None:
        out["compact"]["parity"] = _compact(parity, ["n", "median", "p99", "max_rel", "ref_median", "ref_p99",
                                                     "ref_max_rel", "ok"])
    if cpu_base is not None:
        out["compact"]["cpu"] = round(cpu_base[0])
    del d_src
    return out


def bench_barneshut(args, n, rank, world, local_rank):
    import torch

    import particular_b200 as pb
    dist = _dist_setup(world, local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = pb.CudaContext(local_rank, bh_build=None if args.bh_build == "auto" else args.bh_build)
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
            "counters_last_step_rank0": res["counters_last_step_rank0"], "roofline": res["roofline"]}
    for k in ("cpu_baseline", "parity"):
        if k in res:
            line[k] = res[k]
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
