"""Minimal Barnes-Hut run for ncu captures and regression fingerprints.
Usage (on the GPU box): python scripts/profile_bh.py [N] [iters] [theta] [dist]
Prints phase times, traversal counters and a checksum of the accelerations (a change that is
meant to keep the target groups identical must keep counters and checksum bit-identical)."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from particular_b200._ffi import lib
from tests.conftest import plummer_cloud, uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
THETA = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
DIST = sys.argv[4] if len(sys.argv) > 4 else "plummer"

for kv in os.environ.get("PCUDA_DEBUG", "").split(","):  # e.g. PCUDA_DEBUG=bh_tpl=1,bh_count=0
    if kv:
        k, v = kv.split("=")
        assert lib.pcuda_debug_set(k.encode(), int(v)) == 0, kv

P = plummer_cloud(N) if DIST == "plummer" else uniform_cloud(N)
d_src = torch.from_numpy(P).cuda()
d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
with pb.CudaContext(0, leaf_size=int(os.environ.get("PCUDA_LEAF", "0"))) as ctx:
    bh = pb.BarnesHut(ctx, THETA, pb.Acceleration.checked())
    for it in range(ITERS):
        bh.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
        ctx.sync()
        t = ctx.timings()
        c = bh.last_counters()
        print(f"iter {it}: build {t['build_ms']:.3f} ms traverse {t['compute_ms']:.3f} ms "
              f"launches {t['kernel_launches']} counters {c}")
    out = d_out.cpu().numpy()
    print("finite:", bool(np.isfinite(out).all()), "sha1:", hashlib.sha1(out.tobytes()).hexdigest())
