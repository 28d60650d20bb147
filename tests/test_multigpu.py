"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): torchrun-style world of 2
processes, NCCL all-gather issued by the library, results equal to the single-GPU evaluation."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
import particular_b200 as pb
from tests.conftest import uniform_cloud, plummer_cloud
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = pb.CudaContext(rank)
res = {}
p = uniform_cloud(30001, seed=3)
sh = pb.ShardedBruteForce(ctx, pb.AccelerationSoftened.checked(2.0))
full = sh.compute(p)
single = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(2.0)).compute(p)
res["bf_shape"] = list(full.shape)
res["bf_max_rel"] = float(np.max(np.linalg.norm(full - single, axis=1) / np.linalg.norm(single, axis=1)))
q = plummer_cloud(40003, seed=4)
bh = pb.ShardedBarnesHut(ctx, 0.5, pb.Acceleration.checked(), init_comm=False)
bh.world, bh.rank = sh.world, sh.rank
fullb = bh.compute(q)
singleb = pb.BarnesHut(ctx, 0.5, pb.Acceleration.checked()).compute(q)
truth = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(q)
den = np.linalg.norm(truth, axis=1)
e_sh = np.linalg.norm(fullb - truth, axis=1) / den
e_1 = np.linalg.norm(singleb - truth, axis=1) / den
res["bh_shape"] = list(fullb.shape)
res["bh_err_sharded"] = [float(np.median(e_sh)), float(np.percentile(e_sh, 99)), float(e_sh.max())]
res["bh_err_single"] = [float(np.median(e_1)), float(np.percentile(e_1, 99)), float(e_1.max())]
# partitioned build: one tree per GPU over its key range, walked as a forest (bh_forest = 1 forces it
# on this context; PCUDA_FLAG_BH_PARTITIONED_BUILD / CudaContext(partitioned_build=True) is the API)
from particular_b200 import _ffi
assert _ffi.lib.pcuda_debug_set(b"bh_forest", 1) == 0
assert _ffi.lib.pcuda_debug_set(b"bh_route", 2) == 0   # all-to-all routing (automatic from 4 GPUs on)
fullf = bh.compute(q)
e_f = np.linalg.norm(fullf - truth, axis=1) / den
res["bh_forest_shape"] = list(fullf.shape)
res["bh_err_forest"] = [float(np.median(e_f)), float(np.percentile(e_f, 99)), float(e_f.max())]
res["bh_forest_finite"] = bool(np.isfinite(fullf).all())
bh0 = pb.ShardedBarnesHut(ctx, 0.0, pb.Acceleration.checked(), init_comm=False)
bh0.world, bh0.rank = sh.world, sh.rank
small = uniform_cloud(5001, seed=9)
f0 = bh0.compute(small)
t0 = pb.BruteForce(ctx, pb.Acceleration.checked()).compute(small)
res["bh_forest_theta0_max_rel"] = float(np.max(np.linalg.norm(f0 - t0, axis=1) / np.linalg.norm(t0, axis=1)))
# result routing: rows sent to their owners only (ncclSend / ncclRecv) against the all-gather path
assert _ffi.lib.pcuda_debug_set(b"bh_route", 1) == 0
fullf_ag = bh.compute(q)
assert _ffi.lib.pcuda_debug_set(b"bh_forest", 2) == 0
fullb_ag = bh.compute(q)
assert _ffi.lib.pcuda_debug_set(b"bh_route", 2) == 0
fullb_a2a = bh.compute(q)
assert _ffi.lib.pcuda_debug_set(b"bh_route", 0) == 0
assert _ffi.lib.pcuda_debug_set(b"bh_forest", 0) == 0
res["route_same_forest"] = bool(np.array_equal(fullf, fullf_ag))
res["route_same_replicated"] = bool(np.array_equal(fullb_a2a, fullb_ag) and np.array_equal(fullb, fullb_ag))
r = uniform_cloud(20011, seed=6, massive_ratio=0.01)
sb = pb.ShardedBetween(ctx, pb.AccelerationSoftened.checked(1.0), init_comm=False)
sb.world, sb.rank = sh.world, sh.rank
fulls = sb.compute(pb.Reordered.new(r))
singles = pb.BruteForce(ctx, pb.AccelerationSoftened.checked(1.0)).compute(pb.Reordered.new(r))
res["split_shape"] = list(fulls.shape)
res["split_max_rel"] = float(np.max(np.linalg.norm(fulls - singles, axis=1) / np.linalg.norm(singles, axis=1)))
res["comm_ms"] = ctx.timings()["comm_ms"]
if rank == 0:
    print("RESULT " + json.dumps(res), flush=True)
ctx.close()
dist.destroy_process_group()
'''


def test_two_gpus_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=root)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[7:])
    assert res["bf_shape"] == [30001, 3]
    assert res["bf_max_rel"] <= 1e-5   # same kernel, other source-split boundaries
    assert res["split_shape"] == [20011, 3]
    assert res["split_max_rel"] <= 1e-5
    # identical tree on every GPU; the target groups differ (each rank groups its own block), so
    # the results agree to the theta-approximation error, which must be the same as on one GPU
    assert res["bh_shape"] == [40003, 3]
    for a, b in zip(res["bh_err_sharded"], res["bh_err_single"]):
        assert a <= 1.25 * b + 1e-6, res
    # partitioned build: a forest of per-GPU trees; same error distribution, theta = 0 exact
    assert res["bh_forest_shape"] == [40003, 3] and res["bh_forest_finite"]
    for a, b in zip(res["bh_err_forest"], res["bh_err_single"]):
        assert a <= 1.25 * b + 1e-6, res
    assert res["bh_forest_theta0_max_rel"] <= 2e-5, res
    assert res["route_same_forest"] and res["route_same_replicated"], res
