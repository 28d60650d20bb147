"""Tree-build timing: one-pass (radix) build against the level-wise build it replaces.
Usage (on the GPU box): python scripts/profile_build.py [N] [dist]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from particular_b200._ffi import lib
from tests.conftest import plummer_cloud, uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
DIST = sys.argv[2] if len(sys.argv) > 2 else "plummer"
P = plummer_cloud(N) if DIST == "plummer" else uniform_cloud(N)
d_src = torch.from_numpy(P).cuda()
d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
with pb.CudaContext(0) as ctx:
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    for mode, name in ((1, "level-wise"), (0, "one-pass")):
        assert lib.pcuda_debug_set(b"bh_level_build", mode) == 0
        for it in range(3):
            bh.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
            ctx.sync()
            t = ctx.timings()
        print(f"{name}: N={N} {DIST} build {t['build_ms']:.3f} ms traverse {t['compute_ms']:.3f} ms "
              f"launches {t['kernel_launches']}")
