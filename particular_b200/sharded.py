"""Multi-GPU brute force and Barnes-Hut: one process per GPU, targets sharded, sources all-gathered
over NVLink (Barnes-Hut: the tree is built replicated or partitioned by key range, see ShardedBarnesHut).

New functionality (the reference is single-device).  Each rank owns a
contiguous block of the particle slice (input order is preserved, so concatenating the ranks'
outputs in rank order gives the reference's output order, sequential.rs:101-106).  Per step:

    local {x,y,z,mu} records --(pcuda: pad + in-place ncclAllGather on the context stream)-->
    all records on every GPU --(pair kernel, local targets x all sources)--> local accelerations

PyTorch is plumbing only: device memory for the shards and ``torch.distributed`` to hand rank 0's
ncclUniqueId to the other ranks (and, in ``compute``, to collect the shards' results on the host).
The all-gather on the data path is issued by libparticular_cuda.so on its own communicator and
stream, not by torch.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

__all__ = ["shard_bounds", "shard_capacity", "ShardedBruteForce", "ShardedBarnesHut", "ShardedBetween"]


def shard_capacity(n: int, world: int) -> int:
    """Records per rank slot: ceil(n / world), at least 1 (NCCL needs equal, non-empty slots)."""
    return max(1, -(-n // world))


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the particle slice owned by `rank`."""
    cap = shard_capacity(n, world)
    lo = min(n, rank * cap)
    return lo, min(n, lo + cap)


def _check_tensor(t, cols: int, name: str, device: int):
    """The C ABI reads `t.data_ptr()` as packed float32 rows: anything else would be silently wrong
    (or out of bounds), so it is refused here."""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.device.index != device:
        raise ValueError(f"{name}: tensor lives on cuda:{t.device.index}, the context on cuda:{device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"{name}: expected shape (n, {cols}), got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous (packed rows)")
    if cols == 4 and t.shape[0] and t.data_ptr() % 16:  # {x,y,z,mu} records are read as float4
        raise ValueError(f"{name}: records must start at a 16-byte aligned address")


class _StreamOrder:
    """The library works on its own non-blocking stream.  Entering makes that stream wait for the
    work already queued on torch's current stream (which produced the inputs and allocated the
    outputs); leaving makes torch's current stream wait for the library's work, so that a consumer
    of the returned tensor on torch's stream is ordered after the kernels."""

    def __init__(self, ctx):
        import torch
        self.torch = torch
        self.ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", ctx.device))

    def __enter__(self):
        cur = self.torch.cuda.current_stream(self.ext.device)
        if cur.cuda_stream != self.ext.cuda_stream:
            self.ext.wait_stream(cur)
        return self

    def __exit__(self, *exc):
        cur = self.torch.cuda.current_stream(self.ext.device)
        if cur.cuda_stream != self.ext.cuda_stream:
            cur.wait_stream(self.ext)
        return False


class ShardedBruteForce:
    """``ShardedBruteForce(ctx, interaction).compute(particles)``: the multi-GPU counterpart of
    ``BruteForce(ctx, interaction).compute(particles)`` for the ``&[P]`` storage (all particles
    affect all particles, storage.rs:231-241).  f32 3-D."""

    def __init__(self, ctx, interaction, group=None, init_comm: bool = True):
        import torch.distributed as dist
        self.ctx, self.interaction, self.group = ctx, interaction, group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gathered = None
        self._out = None
        if init_comm and self.world > 1:
            self._init_comm()

    # -- communicator bootstrap: rank 0 makes the ncclUniqueId, torch.distributed carries it --
    def _init_comm(self):
        import torch
        uid = [self.ctx.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(uid, src=0, group=self.group)
        self.ctx.comm_init(uid[0], self.world, self.rank)
        torch.cuda.synchronize()

    # -- device-resident step ------------------------------------------------------------------
    def step_device(self, local, n_total: int):
        """`local`: this rank's (n_local, 4) float32 CUDA tensor of {x,y,z,mu}; `n_total`: particle
        count over all ranks.  Returns the (n_local, 3) CUDA tensor of accelerations (owned by this
        object, overwritten by the next call).  Enqueued on the context stream, ordered after the work
        already queued on torch's current stream; torch's current stream is made to wait for it, so
        the result can be consumed there without a host synchronisation."""
        import torch
        from ._ffi import check, lib
        _check_tensor(local, 4, "local", self.ctx.device)
        cap = shard_capacity(n_total, self.world)
        n_local = int(local.shape[0])
        if self._gathered is None or self._gathered.shape[0] < self.world * cap:
            self._gathered = torch.empty((self.world * cap, 4), dtype=torch.float32,
                                         device=local.device)
        if self._out is None or self._out.shape[0] < max(n_local, 1):
            self._out = torch.empty((max(n_local, 1), 3), dtype=torch.float32, device=local.device)
        it = self.interaction
        with _StreamOrder(self.ctx):
            check(lib.pcuda_bruteforce_f32x3_sharded_dev(
                self.ctx.handle, local.data_ptr(), n_local, cap, it.softening, int(it.is_checked),
                self._gathered.data_ptr(), self._out.data_ptr()), self.ctx.handle)
        return self._out[:n_local]

    # -- host API --------------------------------------------------------------------------------
    def compute_local(self, local_records: np.ndarray, n_total: int,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """Upload this rank's records, run one sharded step, download the local accelerations
        (pcuda_bruteforce_f32x3_sharded: blocking, host buffers)."""
        import ctypes as C

        from ._ffi import check, lib
        it = self.interaction
        if out is None:
            out = np.zeros((len(local_records), 3), dtype=np.float32)
        check(lib.pcuda_bruteforce_f32x3_sharded(
            self.ctx.handle, local_records.ctypes.data_as(C.c_void_p), len(local_records),
            shard_capacity(n_total, self.world), it.softening, int(it.is_checked),
            out.ctypes.data_as(C.c_void_p)), self.ctx.handle)
        return out

    def compute(self, particles, gather: bool = True) -> Optional[np.ndarray]:
        """Every rank passes the same full (n, 4) slice.  Returns all accelerations in slice
        order on every rank (gather=True) or only this rank's block."""
        p = np.ascontiguousarray(particles, dtype=np.float32)
        if p.ndim != 2 or p.shape[1] != 4:
            raise NotImplementedError("sharded brute force is f32 3-D: particles must be (n, 4)")
        n = len(p)
        lo, hi = shard_bounds(n, self.world, self.rank)
        local = self.compute_local(np.ascontiguousarray(p[lo:hi]), n)
        if not gather or self.world == 1:
            return local
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local, group=self.group)
        return np.concatenate(parts, axis=0)


class ShardedBarnesHut(ShardedBruteForce):
    """``ShardedBarnesHut(ctx, theta, interaction).compute(particles)``: the multi-GPU counterpart
    of ``BarnesHut(ctx, theta, interaction).compute(particles)`` for the ``&[P]`` storage.  Every
    rank all-gathers the particle records; the tree is either built whole on every rank
    (replicated build) or, from 4 GPUs on, partitioned by key range — each rank builds the tree of
    its own range, the trees are all-gathered and joined by a small top tree
    (``CudaContext(partitioned_build=...)``).  Each rank walks the tree for the
    targets of its key range and the accelerations are routed back to the ranks that own the
    particles.  f32 3-D."""

    def __init__(self, ctx, theta: float, interaction, group=None, init_comm: bool = True):
        super().__init__(ctx, interaction, group, init_comm)
        self.theta = float(theta)

    def step_device(self, local, n_total: int):
        import torch
        from ._ffi import check, lib
        _check_tensor(local, 4, "local", self.ctx.device)
        cap = shard_capacity(n_total, self.world)
        n_local = int(local.shape[0])
        if self._gathered is None or self._gathered.shape[0] < self.world * cap:
            self._gathered = torch.empty((self.world * cap, 4), dtype=torch.float32,
                                         device=local.device)
        if self._out is None or self._out.shape[0] < max(n_local, 1):
            self._out = torch.empty((max(n_local, 1), 3), dtype=torch.float32, device=local.device)
        it = self.interaction
        with _StreamOrder(self.ctx):
            check(lib.pcuda_barneshut_f32x3_sharded_dev(
                self.ctx.handle, local.data_ptr(), n_local, n_total, self.theta, it.softening,
                int(it.is_checked), self._gathered.data_ptr(), self._out.data_ptr()), self.ctx.handle)
        return self._out[:n_local]

    def compute_local(self, local_records: np.ndarray, n_total: int,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        import ctypes as C

        from ._ffi import check, lib
        it = self.interaction
        if out is None:
            out = np.zeros((len(local_records), 3), dtype=np.float32)
        check(lib.pcuda_barneshut_f32x3_sharded(
            self.ctx.handle, local_records.ctypes.data_as(C.c_void_p), len(local_records), n_total,
            self.theta, it.softening, int(it.is_checked), out.ctypes.data_as(C.c_void_p)),
            self.ctx.handle)
        return out


class ShardedBetween(ShardedBruteForce):
    """``ShardedBetween(ctx, interaction).compute(storage)``: multi-GPU brute force for the storages
    whose affecting set is a (small) subset — ``Between(affected, affecting)``, ``Ordered`` and
    ``Reordered`` (storage.rs:61-95, 153-163, 207-229; BASELINE configs[2]: 10 k massive act on
    16 M massless).  The AFFECTED particles are sharded in contiguous blocks (input order kept, so
    the ranks' outputs concatenate to the reference's output order); the AFFECTING records are
    sharded too and all-gathered over NVLink by the library (16 B each); nothing else is exchanged.
    f32 3-D."""

    def compute_local(self, affected_local: np.ndarray, src_local: np.ndarray, n_src_total: int,
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """`affected_local`: this rank's (n, 3) positions; `src_local`: this rank's block of the
        {x,y,z,mu} affecting records (``shard_bounds(n_src_total, world, rank)``)."""
        import ctypes as C

        from ._ffi import check, lib
        it = self.interaction
        if out is None:
            out = np.zeros((len(affected_local), 3), dtype=np.float32)
        check(lib.pcuda_bruteforce_f32x3_between_sharded(
            self.ctx.handle, affected_local.ctypes.data_as(C.c_void_p), len(affected_local),
            src_local.ctypes.data_as(C.c_void_p), len(src_local),
            shard_capacity(n_src_total, self.world), it.softening, int(it.is_checked),
            out.ctypes.data_as(C.c_void_p)), self.ctx.handle)
        return out

    def step_device(self, affected_local, src_local, n_src_total: int):
        """Device-resident step: `affected_local` (n, 3) and `src_local` (m, 4) float32 CUDA
        tensors.  Returns the (n, 3) accelerations (owned by this object).  Not synchronised."""
        import torch
        from ._ffi import check, lib
        _check_tensor(affected_local, 3, "affected_local", self.ctx.device)
        _check_tensor(src_local, 4, "src_local", self.ctx.device)
        cap = shard_capacity(n_src_total, self.world)
        n = int(affected_local.shape[0])
        if self._gathered is None or self._gathered.shape[0] < self.world * cap:
            self._gathered = torch.empty((self.world * cap, 4), dtype=torch.float32,
                                         device=affected_local.device)
        if self._out is None or self._out.shape[0] < max(n, 1):
            self._out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=affected_local.device)
        it = self.interaction
        with _StreamOrder(self.ctx):
            check(lib.pcuda_bruteforce_f32x3_between_sharded_dev(
                self.ctx.handle, affected_local.data_ptr(), n, src_local.data_ptr(),
                int(src_local.shape[0]), cap, it.softening, int(it.is_checked),
                self._gathered.data_ptr(), self._out.data_ptr()), self.ctx.handle)
        return self._out[:n]

    def compute(self, storage, gather: bool = True) -> Optional[np.ndarray]:
        """Every rank passes the same storage.  Returns all accelerations in the storage's affected
        order on every rank (gather=True) or only this rank's block."""
        from .interface import _resolve
        aff, src = _resolve(storage)
        if src.dtype != np.float32 or src.shape[1] != 4:
            raise NotImplementedError("sharded brute force is f32 3-D")
        if aff is None:
            aff = np.ascontiguousarray(src[:, :3])
        lo, hi = shard_bounds(len(aff), self.world, self.rank)
        slo, shi = shard_bounds(len(src), self.world, self.rank)
        local = self.compute_local(np.ascontiguousarray(aff[lo:hi]),
                                   np.ascontiguousarray(src[slo:shi]), len(src))
        if not gather or self.world == 1:
            return local
        parts = [None] * self.world
        self.dist.all_gather_object(parts, local, group=self.group)
        return np.concatenate(parts, axis=0)
