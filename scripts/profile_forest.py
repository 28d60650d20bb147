"""Key-range-partitioned Barnes-Hut build on ONE GPU (virtual ranks): cost of the partition step and of
one part's build, forest-walk work against the single tree.  Usage: python scripts/profile_forest.py [N]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from particular_b200._ffi import check, lib
from tests.conftest import plummer_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
p = plummer_cloud(N)
d_p = torch.from_numpy(p).cuda()
d_o = torch.empty((N, 3), dtype=torch.float32, device="cuda")
with pb.CudaContext(0) as ctx:
    bh = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked())
    for _ in range(2):
        bh.compute_device(None, N, d_p.data_ptr(), N, d_o.data_ptr())
    ctx.sync()
    t = ctx.timings()
    ref = d_o.clone()
    print(f"single tree: build {t['build_ms']:.3f} ms  traverse {t['compute_ms']:.3f} ms  {bh.last_counters()}")
    for parts in (1, 2, 4, 8, 16):
        for _ in range(2):
            check(lib.pcuda_barneshut_f32x3_partitioned_dev(ctx.handle, d_p.data_ptr(), N, parts, 0.5, 0.0, 1,
                                                            d_o.data_ptr()), ctx.handle)
            ctx.sync()
        t = ctx.timings()
        diff = (d_o - ref).norm(dim=1) / ref.norm(dim=1)
        print(f"parts {parts:2d}: build(all parts) {t['build_ms']:.3f} ms  per part ~{t['build_ms'] / parts:.3f} ms  "
              f"walk(all parts) {t['compute_ms']:.3f} ms  launches {t['kernel_launches']}  "
              f"median |a - a_single| / |a_single| {diff.median().item():.2e}  max {diff.max().item():.2e}")
