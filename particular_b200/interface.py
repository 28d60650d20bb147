"""Host-side mirror of particular's operator interface for the CUDA backend.

Same names, argument meaning and error behaviour as the reference (paths relative to
/root/reference/particular/src):

  Between(affected, affecting)              lib.rs:299-300
  Interaction.compute(storage)              lib.rs:364-370
  Ordered / Reordered / &[P] storages       storage.rs:48-241
  Acceleration / AccelerationSoftened       gravity/newtonian/acceleration.rs:17-58,
                                            gravity/newtonian/acceleration_softened.rs:17-63
  BruteForce(resources.., interaction)      gpu/mod.rs:149-208   (the existing GPU operator)
  BarnesHut(theta, interaction)             sequential.rs:439-543
  RootedOrthtree                            storage.rs:11-46
  GpuCompute extension sugar                gpu/mod.rs:13-37     -> cuda_brute_force / cuda_barnes_hut

A "slice of particles" is a C-contiguous numpy array of shape (n, D+1): one
``GravitationalField {position, m}`` row per particle (gravity/mod.rs:12-18); dtype float32 with
D in {2, 3} or float64 with D == 3.  Results are (n_affected, D) arrays in affected order (the
reference yields an iterator of vectors in the same order).

All arithmetic happens in libparticular_cuda.so; nothing here computes an interaction on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Callable, Optional

import numpy as np

from . import _ffi
from ._ffi import CudaError, check, lib

__all__ = [
    "Between", "Ordered", "Reordered", "Acceleration", "AccelerationSoftened", "CudaContext",
    "BruteForce", "BarnesHut", "RootedOrthtree", "Simulation", "CustomInteraction",
    "check_interaction_source", "cuda_brute_force", "cuda_barnes_hut",
    "is_affecting", "CudaError", "morton_keys",
]


# ---- storages -------------------------------------------------------------------------------------
@dataclass
class Between:
    """``Between(affected, affecting)``: the first is acted upon, the second acts (lib.rs:299-300)."""
    affected: Any
    affecting: Any


def is_affecting(particles: np.ndarray) -> np.ndarray:
    """``GravitationalField::is_affecting`` (gravity/mod.rs:29-34): m != 0, vectorised."""
    return particles[:, -1] != 0


def _as_particles(p) -> np.ndarray:
    a = np.ascontiguousarray(p)
    if a.ndim != 2 or a.shape[1] not in (3, 4) or a.dtype not in (np.float32, np.float64):
        raise TypeError(f"particles must be (n, D+1) float32/float64 with D in (2, 3); got "
                        f"{a.shape} {a.dtype}")
    return a


class Ordered:
    """Affecting particles first (storage.rs:48-138)."""

    def __init__(self, particles: np.ndarray, affecting_len: int):
        self._particles = particles
        self._affecting_len = affecting_len

    @classmethod
    def with_(cls, affecting, non_affecting, is_affecting_fn: Callable = is_affecting):
        """storage.rs:61-80: chain, then affecting_len = first index failing the predicate."""
        a, b = _as_particles(affecting), _as_particles(non_affecting)
        particles = np.ascontiguousarray(np.concatenate([a, b.astype(a.dtype, copy=False)]))
        mask = np.asarray(is_affecting_fn(particles), dtype=bool)
        fails = np.flatnonzero(~mask)
        return cls(particles, int(fails[0]) if len(fails) else len(particles))

    @classmethod
    def new(cls, unordered, is_affecting_fn: Callable = is_affecting):
        """storage.rs:85-95: two stable filter passes."""
        p = _as_particles(unordered)
        mask = np.asarray(is_affecting_fn(p), dtype=bool)
        return cls.with_(p[mask], p[~mask], is_affecting_fn)

    def affecting_len(self) -> int:
        return self._affecting_len

    def affecting(self) -> np.ndarray:
        return self._particles[: self._affecting_len]

    def non_affecting(self) -> np.ndarray:
        return self._particles[self._affecting_len:]

    def particles(self) -> np.ndarray:
        return self._particles


class Reordered:
    """Borrow of the original slice plus an ``Ordered`` copy (storage.rs:141-205)."""

    def __init__(self, unordered, is_affecting_fn: Callable = is_affecting):
        self.unordered = _as_particles(unordered)
        self._ordered = Ordered.new(self.unordered, is_affecting_fn)
        self._fn = is_affecting_fn

    new = classmethod(lambda cls, unordered, fn=is_affecting: cls(unordered, fn))

    def ordered(self) -> Ordered:
        return self._ordered

    def affecting_len(self) -> int:
        return self._ordered.affecting_len()

    def affecting(self) -> np.ndarray:
        return self._ordered.affecting()

    def non_affecting(self) -> np.ndarray:
        return self._ordered.non_affecting()

    def reordered(self) -> np.ndarray:
        return self._ordered.particles()

    def is_affecting_fn(self):
        return self._fn


def _resolve(storage):
    """The storage blanket impls (storage.rs:207-241) -> (affected_positions | None, affecting).
    ``None`` for affected means "the affecting slice itself" (&[P] => Between(slice, slice))."""
    if isinstance(storage, Between):
        aff, src = storage.affected, storage.affecting
        if isinstance(src, RootedOrthtree):
            return aff, src
        src = _as_particles(src)
        if aff is src:
            return None, src
        aff = np.asarray(aff)
        d = src.shape[1] - 1
        if aff.ndim == 1:  # Between(&P1, &[P2]): a single affected particle
            aff = aff[None, :]
        if aff.shape[1] == d + 1:  # affected given as particles: only their positions matter
            aff = aff[:, :d]
        if aff.shape[1] != d:
            raise TypeError(f"affected has dimension {aff.shape[1]}, affecting has {d}")
        return np.ascontiguousarray(aff, dtype=src.dtype), src
    if isinstance(storage, Ordered):  # storage.rs:207-217
        p = storage.particles()
        return np.ascontiguousarray(p[:, :-1]), np.ascontiguousarray(storage.affecting())
    if isinstance(storage, Reordered):  # storage.rs:219-229
        return (np.ascontiguousarray(storage.unordered[:, :-1]),
                np.ascontiguousarray(storage.affecting()))
    return None, _as_particles(storage)  # storage.rs:231-241


# ---- interactions ---------------------------------------------------------------------------------
@dataclass(frozen=True)
class Acceleration:
    """Newtonian acceleration, no softening (acceleration.rs:17-58)."""
    is_checked: bool = True
    softening = 0.0

    @staticmethod
    def checked():
        return Acceleration(True)

    @staticmethod
    def unchecked():
        return Acceleration(False)


@dataclass(frozen=True)
class AccelerationSoftened:
    """Newtonian acceleration with softening (acceleration_softened.rs:17-63)."""
    softening: float = 0.0
    is_checked: bool = True

    @staticmethod
    def checked(softening):
        return AccelerationSoftened(float(softening), True)

    @staticmethod
    def unchecked(softening):
        return AccelerationSoftened(float(softening), False)


def check_interaction_source(source: str) -> str:
    """Compile an interaction source without a device (the analogue of validating a shader);
    returns the compiler log, raises CudaError with the log when it does not compile."""
    log = C.create_string_buffer(1 << 16)
    check(lib.pcuda_interaction_check(source.encode(), log, len(log)))
    return log.value.decode()


class CustomInteraction:
    """A user-defined pair interaction compiled for the device at run time — the counterpart of
    implementing ``InteractionShader<P1, P2>`` (gpu/mod.rs:40-82) for the wgpu operator.

    ``source`` is CUDA C++ defining ``struct Affected``, ``struct Affecting``, ``struct Interaction``,
    ``struct Push`` (push constants) and
    ``__device__ void compute(const Affected &p1, const Affecting &p2, Interaction &out)``.
    ``affected_dtype`` / ``affecting_dtype`` / ``interaction_dtype`` / ``push_dtype`` are numpy
    (structured) dtypes with exactly those layouts (AFFECTED_SIZE / AFFECTING_SIZE /
    INTERACTION_SIZE of the reference trait are checked against the device compiler's sizeof).
    Use it with ``BruteForce(ctx, custom).compute(Between(affected, affecting))``."""

    def __init__(self, ctx: "CudaContext", source: str, affected_dtype, affecting_dtype,
                 interaction_dtype, push_dtype=None, push=None):
        self.ctx = ctx
        self.affected_dtype = np.dtype(affected_dtype)
        self.affecting_dtype = np.dtype(affecting_dtype)
        self.interaction_dtype = np.dtype(interaction_dtype)
        self.push_dtype = None if push_dtype is None else np.dtype(push_dtype)
        self.push = push
        h = C.c_void_p()
        check(lib.pcuda_interaction_create(ctx.handle, source.encode(), C.byref(h)), ctx.handle)
        self._h = h
        sizes = (C.c_uint32 * 4)()
        check(lib.pcuda_interaction_sizes(h, C.byref(sizes)), ctx.handle)
        self.sizes = tuple(int(x) for x in sizes)
        want = (self.affected_dtype.itemsize, self.affecting_dtype.itemsize, self.interaction_dtype.itemsize)
        if want != self.sizes[:3]:
            self.close()
            raise TypeError(f"dtype sizes {want} do not match the device structs {self.sizes[:3]}")
        if self.push_dtype is not None and self.push_dtype.itemsize > self.sizes[3]:
            self.close()
            raise TypeError(f"push dtype has {self.push_dtype.itemsize} bytes, struct Push {self.sizes[3]}")

    def brute_force(self, affected, affecting, push=None) -> np.ndarray:
        a = np.ascontiguousarray(affected, dtype=self.affected_dtype)
        b = np.ascontiguousarray(affecting, dtype=self.affecting_dtype)
        out = np.zeros(len(a), dtype=self.interaction_dtype)
        push = self.push if push is None else push
        pbytes = None
        if push is not None:
            pbytes = np.ascontiguousarray(np.asarray(push, dtype=self.push_dtype).reshape(1))
        check(lib.pcuda_interaction_brute_force(
            self.ctx.handle, self._h, _ptr(a), len(a), _ptr(b), len(b),
            _ptr(pbytes), 0 if pbytes is None else pbytes.nbytes, _ptr(out)), self.ctx.handle)
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self.ctx._h is not None:
            lib.pcuda_interaction_destroy(self.ctx.handle, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- context (the analogue of GpuResources + wgpu::Device + wgpu::Queue, gpu/mod.rs:85-159) -------
class CudaContext:
    """Owns one device, one stream, grow-only buffers and (optionally) an NCCL communicator.
    Create once and reuse across calls ("should not be recreated for every iteration",
    gpu/mod.rs:150-151)."""

    def __init__(self, device: int = 0, leaf_size: int = 0, phase_timings: bool = True,
                 expansion_order: int = 1, partitioned_build: Optional[bool] = None,
                 exact_checked: bool = False, bh_build: Optional[str] = None):
        """phase_timings=False skips the per-phase CUDA events (PCUDA_FLAG_NO_PHASE_TIMINGS):
        about 10 us less per call; timings() then carries only kernel_launches.
        partitioned_build: multi-GPU Barnes-Hut tree build.  True (PCUDA_FLAG_BH_PARTITIONED_BUILD):
        one tree per GPU over its key range, joined by a top tree; False
        (PCUDA_FLAG_BH_REPLICATED_BUILD): every GPU builds the whole tree; None: partitioned from
        4 GPUs on.  bh_build = "let" (PCUDA_FLAG_BH_LET_BUILD) / "partitioned" / "replicated" names the
        build instead: "let" = locally essential trees (particles go to the owners of their key ranges,
        every rank sends the others only what their walks can open), the default from 3 GPUs on when
        every rank gets at least 65536 particles.
        exact_checked (PCUDA_FLAG_EXACT_CHECKED): the f32 brute-force kernels test r^2 == 0 exactly at
        every problem size instead of adding the floor t ~ 1e-19 to r^2 on large problems (see
        include/particular_cuda.h)."""
        flags = (0 if phase_timings else _ffi.FLAG_NO_PHASE_TIMINGS) | \
            (_ffi.FLAG_EXACT_CHECKED if exact_checked else 0) | \
            (0 if partitioned_build is None else
             _ffi.FLAG_BH_PARTITIONED_BUILD if partitioned_build else _ffi.FLAG_BH_REPLICATED_BUILD)
        if bh_build is not None:
            flags |= {"let": _ffi.FLAG_BH_LET_BUILD, "partitioned": _ffi.FLAG_BH_PARTITIONED_BUILD,
                      "replicated": _ffi.FLAG_BH_REPLICATED_BUILD}[bh_build]
        cfg = _ffi.Config(device, flags, leaf_size, expansion_order)
        h = C.c_void_p()
        check(lib.pcuda_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._pinned = []
        self.device = device
        sm, khz = C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        check(lib.pcuda_device_info(h, C.byref(sm), C.byref(khz), name, 128), h)
        self.sm_count, self.sm_clock_khz, self.name = sm.value, khz.value, name.value.decode()

    @property
    def handle(self):
        if self._h is None:
            raise CudaError(_ffi.ERR_NOT_INITIALISED, "context destroyed")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            for p in self._pinned:
                lib.pcuda_host_free(self._h, p)
            self._pinned = []
            lib.pcuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def timings(self) -> dict:
        t = _ffi.Timings()
        check(lib.pcuda_get_timings(self.handle, C.byref(t)), self.handle)
        return t.as_dict()

    def sync(self):
        check(lib.pcuda_sync(self.handle), self.handle)

    @property
    def stream_ptr(self) -> int:
        return int(lib.pcuda_stream(self.handle) or 0)

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """A numpy array backed by page-locked host memory (pcuda_host_alloc): the analogue of
        the mapped staging view the reference packs particles into (gpu/mod.rs:187-195).
        Arrays handed to compute() from here are copied by true asynchronous DMA."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(lib.pcuda_host_alloc(self.handle, max(n, 1), C.byref(p)), self.handle)
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned.append(p)
        return arr

    def probe_fp32(self, packed=True, iters: int = 4096, repeats: int = 5):
        """FP32-pipe microbenchmark; `packed` is a mode 0..7 (see csrc/probe.cu; True = 1)."""
        tf, ms = C.c_double(), C.c_float()
        check(lib.pcuda_probe_fp32(self.handle, int(packed), iters, repeats, C.byref(tf),
                                   C.byref(ms)), self.handle)
        return tf.value, ms.value

    # -- NCCL (multi-GPU, one process per GPU) --
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * _ffi.UNIQUE_ID_BYTES)()
        check(lib.pcuda_comm_unique_id(self.handle, C.byref(buf)), self.handle)
        return bytes(buf)

    def comm_init(self, unique_id: bytes, world_size: int, rank: int):
        buf = (C.c_uint8 * _ffi.UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        check(lib.pcuda_comm_init(self.handle, C.byref(buf), world_size, rank), self.handle)

    @staticmethod
    def comm_init_local(contexts) -> None:
        """In-process communicator for tests (pcuda_comm_init_local): binds the given contexts of this
        process (rank = position in the list; one host thread must drive each) so that the sharded entry
        points run all their ranks on a box with a single GPU."""
        arr = (C.c_void_p * len(contexts))(*[c.handle for c in contexts])
        check(lib.pcuda_comm_init_local(arr, len(contexts)))

    def allgather_dev(self, send_ptr: int, recv_ptr: int, bytes_per_rank: int):
        check(lib.pcuda_comm_allgather_dev(self.handle, send_ptr, recv_ptr, bytes_per_rank),
              self.handle)


def _suffix(src: np.ndarray) -> str:
    d = src.shape[1] - 1
    key = (src.dtype.type, d)
    table = {(np.float32, 3): "f32x3", (np.float32, 2): "f32x2", (np.float64, 3): "f64x3",
             (np.float64, 2): "f64x2"}
    if key not in table:
        # the reference does the same for shader dimensions it lacks: unimplemented!()
        # (gravity/impls/mod.rs:362, 374)
        raise NotImplementedError(f"no CUDA kernel for {src.dtype} in {d} dimensions")
    return table[key]


def _out_array(out, shape, dtype) -> np.ndarray:
    if out is None:
        return np.zeros(shape, dtype=dtype)
    if out.shape != shape or out.dtype != dtype or not out.flags.c_contiguous:
        raise TypeError(f"out must be a C-contiguous {shape} {np.dtype(dtype)} array")
    return out


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RootedOrthtree:
    """A tree built on the device over the affecting particles (storage.rs:11-46).  f32 only."""

    def __init__(self, ctx: CudaContext, particles):
        p = _as_particles(particles)
        if p.dtype != np.float32:
            raise NotImplementedError("device trees are f32")
        self.ctx, self.dim, self.n = ctx, p.shape[1] - 1, len(p)
        h = C.c_void_p()
        check(lib.pcuda_tree_build_f32(ctx.handle, self.dim, _ptr(p), len(p), C.byref(h)),
              ctx.handle)
        self._h = h
        info = _ffi.TreeInfo()
        check(lib.pcuda_tree_info_get(h, C.byref(info)), ctx.handle)
        self.info = info
        self.n_nodes, self.n_levels = int(info.n_nodes), int(info.n_levels)

    new = classmethod(lambda cls, ctx, particles: cls(ctx, particles))

    def read(self, which: int) -> np.ndarray:
        n, m, d = self.n, self.n_nodes, self.dim
        shape, dt = {
            _ffi.TREE_KEYS: ((n,), np.uint64), _ffi.TREE_PERM: ((n,), np.uint32),
            _ffi.TREE_NODE_BEGIN: ((m,), np.uint32), _ffi.TREE_NODE_COUNT: ((m,), np.uint32),
            _ffi.TREE_NODE_LEVEL: ((m,), np.uint32), _ffi.TREE_NODE_FIRST_CHILD: ((m,), np.uint32),
            _ffi.TREE_NODE_NUM_CHILDREN: ((m,), np.uint32),
            _ffi.TREE_NODE_COM_MASS: ((m, d + 1), np.float32)}[which]
        out = np.zeros(shape, dtype=dt)
        check(lib.pcuda_tree_read(self.ctx.handle, self._h, which, _ptr(out), out.nbytes),
              self.ctx.handle)
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self.ctx._h is not None:
            lib.pcuda_tree_destroy(self.ctx.handle, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def morton_keys(ctx: CudaContext, particles):
    """Sorted Morton keys, stable sort permutation and the root cube of a particle slice
    (pcuda_morton_f32x3 / _f32x2: the first half of the tree build).  Returns
    ``(keys uint64[n], perm uint32[n], TreeInfo)``."""
    p = _as_particles(particles)
    if p.dtype != np.float32:
        raise NotImplementedError("device trees are f32")
    d = p.shape[1] - 1
    keys = np.zeros(len(p), dtype=np.uint64)
    perm = np.zeros(len(p), dtype=np.uint32)
    info = _ffi.TreeInfo()
    fn = lib.pcuda_morton_f32x3 if d == 3 else lib.pcuda_morton_f32x2
    check(fn(ctx.handle, _ptr(p), len(p), _ptr(keys), _ptr(perm), C.byref(info)), ctx.handle)
    return keys, perm, info


# ---- algorithms -----------------------------------------------------------------------------------
class BruteForce:
    """Brute-force algorithm on the GPU (the CUDA counterpart of gpu::BruteForce,
    gpu/mod.rs:149-208): ``BruteForce(ctx, interaction).compute(storage)``."""

    def __init__(self, ctx: CudaContext, interaction):
        self.ctx, self.interaction = ctx, interaction

    def compute(self, storage, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out`: optional preallocated (n_affected, D) result array (e.g. from
        ``ctx.pinned_empty``); by default a fresh array is returned, as the reference returns a
        fresh Vec (gpu/mod.rs:184, 205)."""
        if isinstance(self.interaction, CustomInteraction):
            if isinstance(storage, Between):
                return self.interaction.brute_force(storage.affected, storage.affecting)
            if self.interaction.affected_dtype != self.interaction.affecting_dtype:
                raise TypeError("a slice storage needs Affected and Affecting to be the same type; "
                                "use Between(affected, affecting)")
            return self.interaction.brute_force(storage, storage)  # &[P] => Between(slice, slice)
        aff, src = _resolve(storage)
        sfx = _suffix(src)
        d = src.shape[1] - 1
        na = len(src) if aff is None else len(aff)
        out = _out_array(out, (na, d), src.dtype)
        fn = getattr(lib, f"pcuda_bruteforce_{sfx}")
        check(fn(self.ctx.handle, _ptr(aff), na, _ptr(src), len(src),
                 self.interaction.softening, int(self.interaction.is_checked), _ptr(out)),
              self.ctx.handle)
        return out

    def compute_device(self, affected_ptr: Optional[int], n_affected: int, affecting_ptr: int,
                       n_affecting: int, out_ptr: int, suffix: str = "f32x3") -> None:
        """Device-resident variant: raw device pointers, enqueued on the context stream."""
        fn = getattr(lib, f"pcuda_bruteforce_{suffix}_dev")
        check(fn(self.ctx.handle, affected_ptr, n_affected, affecting_ptr, n_affecting,
                 self.interaction.softening, int(self.interaction.is_checked), out_ptr),
              self.ctx.handle)


class BarnesHut:
    """Barnes-Hut on the GPU: ``BarnesHut(ctx, theta, interaction).compute(storage)``
    (sequential.rs:439-543 semantics: the tree is rebuilt on every call unless the storage is
    ``Between(affected, RootedOrthtree)``)."""

    def __init__(self, ctx: CudaContext, theta: float, interaction):
        self.ctx, self.theta, self.interaction = ctx, float(theta), interaction

    new = classmethod(lambda cls, ctx, theta, interaction: cls(ctx, theta, interaction))

    def compute(self, storage, out: Optional[np.ndarray] = None) -> np.ndarray:
        aff, src = _resolve(storage)
        it = self.interaction
        if isinstance(src, RootedOrthtree):
            aff = np.asarray(aff)
            if aff.ndim == 1:
                aff = aff[None, :]
            if aff.shape[1] == src.dim + 1:
                aff = aff[:, : src.dim]
            aff = np.ascontiguousarray(aff, dtype=np.float32)
            out = _out_array(out, (len(aff), src.dim), np.float32)
            check(lib.pcuda_tree_traverse_f32(self.ctx.handle, src._h, _ptr(aff), len(aff),
                                              self.theta, it.softening, int(it.is_checked),
                                              _ptr(out)), self.ctx.handle)
            return out
        sfx = _suffix(src)
        d = src.shape[1] - 1
        na = len(src) if aff is None else len(aff)
        out = _out_array(out, (na, d), src.dtype)
        fn = getattr(lib, f"pcuda_barneshut_{sfx}")
        check(fn(self.ctx.handle, _ptr(aff), na, _ptr(src), len(src), self.theta, it.softening,
                 int(it.is_checked), _ptr(out)), self.ctx.handle)
        return out

    def compute_partitioned(self, particles, parts: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """The key-range-partitioned build of the multi-GPU path run on this one GPU, part after
        part ("virtual ranks"): a forest of `parts` trees over the same root cube, walked for all
        particles (pcuda_barneshut_f32x3_partitioned).  f32 3-D, `&[P]` storage."""
        p = np.ascontiguousarray(particles, dtype=np.float32)
        if p.ndim != 2 or p.shape[1] != 4:
            raise NotImplementedError("the partitioned build is f32 3-D: particles must be (n, 4)")
        it = self.interaction
        out = _out_array(out, (len(p), 3), np.float32)
        check(lib.pcuda_barneshut_f32x3_partitioned(self.ctx.handle, _ptr(p), len(p), int(parts),
                                                    self.theta, it.softening, int(it.is_checked),
                                                    _ptr(out)), self.ctx.handle)
        return out

    def compute_device(self, affected_ptr: Optional[int], n_affected: int, affecting_ptr: int,
                       n_affecting: int, out_ptr: int, suffix: str = "f32x3") -> None:
        it = self.interaction
        fn = getattr(lib, f"pcuda_barneshut_{suffix}_dev")
        check(fn(self.ctx.handle, affected_ptr, n_affected, affecting_ptr, n_affecting, self.theta,
                 it.softening, int(it.is_checked), out_ptr), self.ctx.handle)

    def last_counters(self) -> dict:
        c = (C.c_uint64 * 5)()
        check(lib.pcuda_tree_last_counters(self.ctx.handle, C.byref(c)), self.ctx.handle)
        return {"node_interactions": int(c[0]), "particle_interactions": int(c[1]),
                "node_tests": int(c[2]), "list_entries": int(c[3]), "groups": int(c[4])}


# ---- device-resident stepping -------------------------------------------------
class Simulation:
    """Particles, velocities and accelerations resident on the device; ``step(n)`` runs n times

        accelerations = algorithm.compute(storage)           # BruteForce or BarnesHut
        velocity += acceleration * dt; position += velocity * dt

    i.e. the loop every caller of the reference writes around ``compute`` (examples/simple/src/
    main.rs:45-59; circular_orbit!, gravity/newtonian/mod.rs:318-331) without the per-call
    upload / read-back of the wgpu operator (gpu/resources.rs:37-39, 318-349).

    ``algorithm`` is a ``BruteForce`` or ``BarnesHut`` object (it supplies the context, the
    interaction and theta).  ``affecting="massive"`` is the ``Reordered`` storage: every particle is
    affected, only those with mu != 0 affect (storage.rs:153-163, 219-229)."""

    def __init__(self, algorithm, particles, velocities=None, *, dt: float, affecting: str = "all",
                 graph: bool = True):
        p = _as_particles(particles)
        _suffix(p)
        d = p.shape[1] - 1
        if affecting not in ("all", "massive"):
            raise ValueError("affecting must be 'all' or 'massive'")
        if velocities is not None:
            velocities = np.ascontiguousarray(velocities, dtype=p.dtype)
            if velocities.shape != (len(p), d):
                raise TypeError(f"velocities must be {(len(p), d)}, got {velocities.shape}")
        self.ctx, self.algorithm = algorithm.ctx, algorithm
        self.n, self.dim, self.dtype = len(p), d, p.dtype
        self._flags = ((_ffi.SIM_AFFECTING_MASSIVE_ONLY if affecting == "massive" else 0)
                       | (0 if graph else _ffi.SIM_NO_GRAPH))
        self.dt = float(dt)
        h = C.c_void_p()
        cfg = self._config()
        check(lib.pcuda_sim_create(self.ctx.handle, C.byref(cfg), _ptr(p), _ptr(velocities), len(p),
                                   C.byref(h)), self.ctx.handle)
        self._h = h

    def _config(self) -> "_ffi.SimConfig":
        alg, it = self.algorithm, self.algorithm.interaction
        is_bh = isinstance(alg, BarnesHut)
        return _ffi.SimConfig(self.dim, _ffi.F64 if self.dtype == np.float64 else _ffi.F32,
                              _ffi.BARNES_HUT if is_bh else _ffi.BRUTE_FORCE, self._flags,
                              alg.theta if is_bh else 0.0, float(it.softening), self.dt,
                              int(it.is_checked), 0)

    def configure(self, algorithm=None, dt: Optional[float] = None) -> None:
        """Swap the algorithm / interaction / dt of a live simulation."""
        if algorithm is not None:
            self.algorithm = algorithm
        if dt is not None:
            self.dt = float(dt)
        cfg = self._config()
        check(lib.pcuda_sim_configure(self.ctx.handle, self._h, C.byref(cfg)), self.ctx.handle)

    def step(self, n_steps: int = 1) -> None:
        """Enqueue n_steps steps on the context stream (asynchronous; read() waits)."""
        check(lib.pcuda_sim_step(self.ctx.handle, self._h, int(n_steps)), self.ctx.handle)

    def read(self, particles: bool = True, velocities: bool = True, accelerations: bool = False):
        """Blocking read-back -> (particles, velocities, accelerations); skipped ones are None."""
        p = np.empty((self.n, self.dim + 1), self.dtype) if particles else None
        v = np.empty((self.n, self.dim), self.dtype) if velocities else None
        a = np.empty((self.n, self.dim), self.dtype) if accelerations else None
        check(lib.pcuda_sim_read(self.ctx.handle, self._h, _ptr(p), _ptr(v), _ptr(a)),
              self.ctx.handle)
        return p, v, a

    def particles(self) -> np.ndarray:
        return self.read(True, False, False)[0]

    def velocities(self) -> np.ndarray:
        return self.read(False, True, False)[1]

    def accelerations(self) -> np.ndarray:
        return self.read(False, False, True)[2]

    def info(self) -> dict:
        i = _ffi.SimInfo()
        check(lib.pcuda_sim_info(self._h, C.byref(i)), self.ctx.handle)
        return {k: getattr(i, k) for k, _ in i._fields_}

    def close(self):
        if getattr(self, "_h", None) is not None and self.ctx._h is not None:
            lib.pcuda_sim_destroy(self.ctx.handle, self._h)
        self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- extension-trait sugar (GpuCompute, gpu/mod.rs:13-37) -------------------------------------------
def cuda_brute_force(storage, ctx: CudaContext, interaction) -> np.ndarray:
    return BruteForce(ctx, interaction).compute(storage)


def cuda_barnes_hut(storage, ctx: CudaContext, theta: float, interaction) -> np.ndarray:
    return BarnesHut(ctx, theta, interaction).compute(storage)
