"""Single-GPU harness for the multi-GPU paths: `world` contexts on one device, bound by the library's
in-process communicator (pcuda_comm_init_local), one host thread per rank.  Every collective of the
sharded entry points then is a rendezvous of the threads and a set of device-to-device copies, so the
sharding, the tree exchange and the result routing run exactly the code of a multi-GPU job — only the
transport is not NCCL.  (tests/test_multigpu.py runs the same paths over NCCL when the box has the GPUs.)"""
import threading

import numpy as np


class LocalWorld:
    def __init__(self, world, device=0, **ctx_kw):
        import particular_b200 as pb
        self.pb, self.world = pb, world
        self.ctxs = [pb.CudaContext(device, **ctx_kw) for _ in range(world)]
        pb.CudaContext.comm_init_local(self.ctxs)

    def close(self):
        for c in self.ctxs:
            c.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def run(self, fn):
        """fn(rank, ctx) on one thread per rank; returns the list of results; re-raises the first error
        after every thread has ended (a rank that fails before a collective would leave the others
        waiting, so errors inside the library are returned by it after the rendezvous)."""
        out, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                out[r] = fn(r, self.ctxs[r])
            except BaseException as e:  # noqa: BLE001
                err[r] = e

        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        errs = [(r, e) for r, e in enumerate(err) if e is not None]
        if errs:  # the rank that failed first left the others waiting at the next collective: show its error
            real = [x for x in errs if "did not reach the collective" not in str(x[1])] or errs
            raise RuntimeError("; ".join(f"rank {r}: {e}" for r, e in real)) from real[0][1]
        return out

    # ---- the sharded operators over this world ------------------------------------------------------
    def barnes_hut(self, particles, theta, interaction=None, n_total=None):
        """ShardedBarnesHut.compute_local on every rank; returns the accelerations in input order."""
        pb = self.pb
        p = np.ascontiguousarray(particles, np.float32)
        n = len(p)
        interaction = interaction or pb.Acceleration.checked()

        def rank_fn(r, ctx):
            bh = pb.ShardedBarnesHut(ctx, theta, interaction, init_comm=False)
            bh.world, bh.rank = self.world, r
            lo, hi = pb.shard_bounds(n, self.world, r)
            return bh.compute_local(np.ascontiguousarray(p[lo:hi]), n)

        return np.concatenate(self.run(rank_fn), axis=0)

    def brute_force(self, particles, interaction=None):
        pb = self.pb
        p = np.ascontiguousarray(particles, np.float32)
        n = len(p)
        interaction = interaction or pb.Acceleration.checked()

        def rank_fn(r, ctx):
            sh = pb.ShardedBruteForce(ctx, interaction, init_comm=False)
            sh.world, sh.rank = self.world, r
            lo, hi = pb.shard_bounds(n, self.world, r)
            return sh.compute_local(np.ascontiguousarray(p[lo:hi]), n)

        return np.concatenate(self.run(rank_fn), axis=0)

    def between(self, storage, interaction):
        from particular_b200.interface import _resolve
        pb = self.pb
        aff, src = _resolve(storage)
        if aff is None:
            aff = np.ascontiguousarray(src[:, :3])

        def rank_fn(r, ctx):
            sb = pb.ShardedBetween(ctx, interaction, init_comm=False)
            sb.world, sb.rank = self.world, r
            lo, hi = pb.shard_bounds(len(aff), self.world, r)
            slo, shi = pb.shard_bounds(len(src), self.world, r)
            return sb.compute_local(np.ascontiguousarray(aff[lo:hi]), np.ascontiguousarray(src[slo:shi]), len(src))

        return np.concatenate(self.run(rank_fn), axis=0)
