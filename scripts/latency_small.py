"""One-shot host-API latency at small N (the reference's criterion sizes): pageable vs pinned host
buffers, brute force and Barnes-Hut.  Usage (GPU box): python scripts/latency_small.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particular_b200 as pb
from tests.conftest import uniform_cloud


def bench(fn, reps=400):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return 1e6 * (time.perf_counter() - t0) / reps


with pb.CudaContext(0) as ctx:
    for n in (64, 1024, 4096, 16384, 65536):
        p = uniform_cloud(n)
        bf = pb.BruteForce(ctx, pb.Acceleration.checked())
        bh = pb.BarnesHut(ctx, 0.7, pb.Acceleration.checked())
        hp = ctx.pinned_empty((n, 4), np.float32)
        hp[:] = p
        ho = ctx.pinned_empty((n, 3), np.float32)
        out = np.zeros((n, 3), np.float32)
        t_page = bench(lambda: bf.compute(p, out=out))
        t_pin = bench(lambda: bf.compute(hp, out=ho))
        tm = ctx.timings()
        t_bh = bench(lambda: bh.compute(hp, out=ho), reps=100)
        tb = ctx.timings()
        print(f"N={n:6d}: brute force pageable {t_page:8.1f} us  pinned {t_pin:8.1f} us  "
              f"(device phases: up {1e3*tm['upload_ms']:.1f} compute {1e3*tm['compute_ms']:.1f} "
              f"down {1e3*tm['download_ms']:.1f} us, {tm['kernel_launches']} launches);  "
              f"barnes-hut(0.7) pinned {t_bh:8.1f} us (build {1e3*tb['build_ms']:.0f} traverse "
              f"{1e3*tb['compute_ms']:.0f} us, {tb['kernel_launches']} launches)")
