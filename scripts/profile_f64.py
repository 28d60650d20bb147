"""f64 brute force (BASELINE configs[4]) for ncu captures.  Usage: python scripts/profile_f64.py [N] [iters]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from tests.conftest import uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262_144
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = uniform_cloud(N).astype(np.float64)
d_p = torch.from_numpy(p).cuda()
d_o = torch.empty((N, 3), dtype=torch.float64, device="cuda")
with pb.CudaContext(0) as ctx:
    bf = pb.BruteForce(ctx, pb.Acceleration.checked())
    for it in range(ITERS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.ExternalStream(ctx.stream_ptr)
        e0.record(s)
        bf.compute_device(None, N, d_p.data_ptr(), N, d_o.data_ptr(), "f64x3")
        e1.record(s)
        ctx.sync()
        ms = e0.elapsed_time(e1)
        print(f"iter {it}: {ms:.3f} ms  {N * N / ms / 1e6:.1f} Gpairs/s  {20 * N * N / ms / 1e9:.2f} TFLOP/s (20 flop/pair)")
    import hashlib
    print("sha1", hashlib.sha1(d_o.cpu().numpy().tobytes()).hexdigest())
