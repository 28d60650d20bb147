"""GPU tuning run: brute-force correctness spot-check + throughput of each kernel variant.
Usage (on the GPU box): python scripts/tune_bf.py [N]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import oracle
import particular_b200 as pb
from particular_b200._ffi import lib
from tests.conftest import rel_err, uniform_cloud

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = pb.CudaContext(0)
print("device", ctx.name, ctx.sm_count, "SMs", ctx.sm_clock_khz, "kHz")
for packed in (False, True):
    tf, ms = ctx.probe_fp32(packed, 8192, 5)
    print(f"fp32 probe packed={packed}: {tf:.2f} TFLOP/s ({ms:.3f} ms)")

# correctness spot check
P = uniform_cloud(3000)
ref = oracle.brute_force(P[:, :3], P)
for tp in (1, 2, 4):
    lib.pcuda_debug_set(b"bf_tp", tp)
    got = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(P)
    print("tp", tp, "max rel err vs oracle", rel_err(got, ref).max())
    got = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(1.5)).compute(P)
    refs = oracle.brute_force(P[:, :3], P, 1.5)
    print("tp", tp, "softened max rel err", rel_err(got, refs).max())

P = uniform_cloud(N)
d_src = torch.from_numpy(P).cuda()
d_out = torch.zeros((N, 3), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
res = {}
for eps, name, clamp, waves in ((0.0, "checked_additive_w16", 3, 16), (0.0, "checked_additive_w48", 3, 48),
                                (0.0, "checked_additive_w96", 3, 96), (0.0, "checked_fmnmx_w48", 2, 48),
                                (0.0, "checked_select_w48", 1, 48), (1.0, "softened_w48", 1, 48)):
    lib.pcuda_debug_set(b"bf_clamp", clamp)
    lib.pcuda_debug_set(b"bf_waves", waves)
    inter = pb.AccelerationSoftened.checked(eps) if eps else pb.Acceleration.checked()
    bf = pb.BruteForce(ctx, inter)
    for tp in (4, 2):
        lib.pcuda_debug_set(b"bf_tp", tp)
        for _ in range(2):
            bf.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            bf.compute_device(None, N, d_src.data_ptr(), N, d_out.data_ptr())
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / reps
        gp = N * N / ms / 1e6
        res[f"{name}_tp{tp}"] = gp
        print(f"{name} tp={tp}: {ms:.2f} ms  {gp:.1f} Gpairs/s  {gp*20/1e3:.2f} TFLOP/s(20/pair)")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/tune_bf.json", "w"))
