"""Brute force with few targets and many sources (what one GPU of an 8-GPU run sees: 125k targets x 1M
sources): throughput of each targets-per-thread variant against the automatic choice.
Usage (on the GPU box): python scripts/tune_bf_rect.py"""
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
for na in (2048, 16384, 62500, 125000, 250000, 500000, 1000000):
    d_out = torch.zeros((na, 3), dtype=torch.float32, device="cuda")
    line = f"na={na:8d} nb={NB}:"
    for tp in (0, 1, 2, 4):
        lib.pcuda_debug_set(b"bf_tp", tp)
        tgt = None if na == NB else d_src[:na, :3].contiguous()
        t_ptr = None if tgt is None else tgt.data_ptr()
        for _ in range(2):
            bf.compute_device(t_ptr, na, d_src.data_ptr(), NB, d_out.data_ptr())
        ctx.sync()
        reps = max(2, int(2e11 / (na * NB)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            bf.compute_device(t_ptr, na, d_src.data_ptr(), NB, d_out.data_ptr())
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        line += f"  tp={tp}: {ms:8.3f} ms frac {20.0 * na * NB / (ms * 1e-3) / 1e12 / peak:.4f}"
    print(line, flush=True)
lib.pcuda_debug_set(b"bf_tp", 0)
