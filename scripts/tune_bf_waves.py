"""Brute force, 125k targets x 1M sources (one GPU's share at 8 GPUs): waves of equal-cost CTAs per slot.
Usage (on the GPU box): python scripts/tune_bf_waves.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particular_b200 as pb
from particular_b200._ffi import lib
from tests.conftest import uniform_cloud

NB = 1_000_000
ctx = pb.CudaContext(0)
P = uniform_cloud(NB)
d_src = torch.from_numpy(P).cuda()
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
bf = pb.BruteForce(ctx, pb.Acceleration.checked())
peak = ctx.sm_count * 128 * 2 * ctx.sm_clock_khz * 1e3 / 1e12
for na in (125000, 250000, 500000):
    tgt = d_src[:na, :3].contiguous()
    d_out = torch.zeros((na, 3), dtype=torch.float32, device="cuda")
    for tp in (4, 2):
        lib.pcuda_debug_set(b"bf_tp", tp)
        line = f"na={na} tp={tp}:"
        for waves in (8, 12, 16, 24, 32, 40, 48, 64, 96):
            lib.pcuda_debug_set(b"bf_waves", waves)
            for _ in range(2):
                bf.compute_device(tgt.data_ptr(), na, d_src.data_ptr(), NB, d_out.data_ptr())
            ctx.sync()
            reps = 4
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                bf.compute_device(tgt.data_ptr(), na, d_src.data_ptr(), NB, d_out.data_ptr())
            e1.record(stream)
            ctx.sync()
            ms = e0.elapsed_time(e1) / reps
            line += f" w{waves}: {20.0 * na * NB / (ms * 1e-3) / 1e12 / peak:.4f}"
        print(line, flush=True)
lib.pcuda_debug_set(b"bf_tp", 0)
lib.pcuda_debug_set(b"bf_waves", 48)
