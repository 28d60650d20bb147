"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo process group.

The GPU step itself (pcuda_bruteforce_f32x3_sharded) needs B200s; here the exchange + local
evaluation is replaced by a gloo all-gather + the CPU oracle, so that shard bounds, padding,
rank-order concatenation and the communicator bootstrap plumbing are exercised end to end."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import uniform_cloud


def test_shard_bounds_cover_and_preserve_order():
    from particular_b200.sharded import shard_bounds, shard_capacity
    for n in (0, 1, 2, 7, 8, 9, 1000, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            cap = shard_capacity(n, world)
            assert cap >= 1 and cap * world >= n
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            assert all(b - a <= cap for a, b in spans)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from particular_b200.sharded import ShardedBruteForce, shard_capacity

        class GlooOracleSharded(ShardedBruteForce):
            """Stand-in for the device step: same slot layout (capacity, zero-mass padding far
            away), exchange over gloo, evaluation by the CPU oracle."""

            def compute_local(self, local_records, n_total, out=None):
                cap = shard_capacity(n_total, self.world)
                slot = np.zeros((cap, 4), np.float32)
                slot[:, :3] = 1e18
                slot[: len(local_records)] = local_records
                parts = [torch.empty((cap, 4)) for _ in range(self.world)]
                self.dist.all_gather(parts, torch.from_numpy(slot))
                gathered = torch.cat(parts).numpy()
                return oracle.brute_force(local_records[:, :3], gathered)

        class FakeCtx:
            def comm_unique_id(self):
                return bytes(range(128))

            def comm_init(self, uid, world, rank):
                self.uid, self.world, self.rank = uid, world, rank

        ctx = FakeCtx()
        sh = GlooOracleSharded(ctx, None, init_comm=False)
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        assert uid[0] == bytes(range(128))
        p = uniform_cloud(n)
        full = sh.compute(p)
        ref = oracle.brute_force(p[:, :3], p)
        q.put((rank, full.shape == ref.shape and bool(np.array_equal(full, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 37, 256])
def test_world2_gloo_sharded_equals_single(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(ok for _, ok in res), res


def _worker_between(rank, world, port, n_massive, n_massless, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from particular_b200.interface import Reordered
        from particular_b200.sharded import ShardedBetween, shard_capacity

        class GlooOracleBetween(ShardedBetween):
            """Stand-in for pcuda_bruteforce_f32x3_between_sharded: same source-slot layout,
            exchange over gloo, evaluation by the CPU oracle."""

            def compute_local(self, affected_local, src_local, n_src_total, out=None):
                cap = shard_capacity(n_src_total, self.world)
                slot = np.zeros((cap, 4), np.float32)
                slot[:, :3] = 1e18
                slot[: len(src_local)] = src_local
                parts = [torch.empty((cap, 4)) for _ in range(self.world)]
                self.dist.all_gather(parts, torch.from_numpy(slot))
                gathered = torch.cat(parts).numpy()
                if len(affected_local) == 0:
                    return np.zeros((0, 3), np.float32)
                return oracle.brute_force(affected_local, gathered)

        p = uniform_cloud(n_massive + n_massless, seed=5)
        massless = np.random.default_rng(9).permutation(len(p))[:n_massless]
        p[massless, 3] = 0.0
        storage = Reordered(p)  # affected: all, input order; affecting: the massive, input order
        sh = GlooOracleBetween(None, None, init_comm=False)
        full = sh.compute(storage)
        massive = p[p[:, 3] != 0]
        ref = oracle.brute_force(p[:, :3], massive) if len(massive) else np.zeros((len(p), 3), np.float32)
        q.put((rank, full.shape == ref.shape and bool(np.array_equal(full, ref))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_massive,n_massless", [(1, 5), (7, 130), (33, 0), (2, 1)])
def test_world2_gloo_between_equals_single(n_massive, n_massless):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_between, args=(r, world, port, n_massive, n_massless, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(ok for _, ok in res), res


def test_device_step_refuses_tensors_the_c_abi_would_misread():
    """ShardedBruteForce / ShardedBarnesHut / ShardedBetween.step_device hand `data_ptr()` to the C ABI,
    which reads packed float32 rows: host tensors are refused before any pointer is taken (the dtype /
    shape / contiguity / device checks need a GPU: tests/test_properties_gpu.py)."""
    import pytest
    import torch
    from particular_b200.sharded import _check_tensor
    with pytest.raises(TypeError):
        _check_tensor(torch.zeros(8, 4), 4, "local", 0)
    with pytest.raises(TypeError):
        _check_tensor([[0.0] * 4], 4, "local", 0)
