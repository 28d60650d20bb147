"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  ``particular_b200`` never imports it.

Array conventions (all C-contiguous numpy arrays):
  affected   (na, D)    positions
  affecting  (nb, D+1)  {position, mu}   — ``GravitationalField`` (gravity/mod.rs:12-18)
  result     (na, D)    accelerations in affected order
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_SFX = {(np.dtype(np.float32), 3): "f32x3", (np.dtype(np.float32), 2): "f32x2",
        (np.dtype(np.float64), 3): "f64x3", (np.dtype(np.float64), 2): "f64x2"}


def build(force: bool = False) -> str:
    """Compile liboracle.so with oracle/Makefile (gcc)."""
    srcs = [os.path.join(_HERE, f) for f in ("particular_oracle.c", "oracle_impl.inc", "oracle_octree.inc",
                                              "baseline_simd.c", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.oracle_abi_version() == 1
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _prep(affected, affecting):
    affecting = np.ascontiguousarray(affecting)
    dt = affecting.dtype
    d = affecting.shape[1] - 1
    affected = np.ascontiguousarray(affected, dtype=dt)
    assert affected.ndim == 2 and affected.shape[1] == d, (affected.shape, affecting.shape)
    return affected, affecting, dt, d, _SFX[(dt, d)]


def _scalar(dt, v):
    return C.c_float(v) if dt == np.float32 else C.c_double(v)


def brute_force(affected, affecting, softening=0.0, checked=True):
    """sequential::BruteForce over Between(affected, affecting) (sequential.rs:178-209)."""
    affected, affecting, dt, d, sfx = _prep(affected, affecting)
    out = np.zeros((affected.shape[0], d), dtype=dt)
    getattr(lib(), f"oracle_bruteforce_{sfx}")(
        _ptr(affected), C.c_size_t(len(affected)), _ptr(affecting), C.c_size_t(len(affecting)),
        _scalar(dt, softening), C.c_int(int(checked)), _ptr(out))
    return out


def brute_force_parallel(affected, affecting, softening=0.0, checked=True):
    """parallel::BruteForce (parallel.rs:195-232); identical results, OpenMP threads."""
    affected, affecting, dt, d, sfx = _prep(affected, affecting)
    out = np.zeros((affected.shape[0], d), dtype=dt)
    getattr(lib(), f"oracle_bruteforce_parallel_{sfx}")(
        _ptr(affected), C.c_size_t(len(affected)), _ptr(affecting), C.c_size_t(len(affecting)),
        _scalar(dt, softening), C.c_int(int(checked)), _ptr(out))
    return out


def brute_force_exact(affected, affecting, softening=0.0, checked=True):
    """The same sum in extended precision (float64 result): the 'truth' for error comparisons."""
    affected, affecting, dt, d, sfx = _prep(affected, affecting)
    out = np.zeros((affected.shape[0], d), dtype=np.float64)
    getattr(lib(), f"oracle_bruteforce_exact_{sfx}")(
        _ptr(affected), C.c_size_t(len(affected)), _ptr(affecting), C.c_size_t(len(affecting)),
        _scalar(dt, softening), C.c_int(int(checked)), _ptr(out))
    return out


def brute_force_abs(affected, affecting, softening=0.0):
    """S_i = sum_j |term_ij| in extended precision: S_i / |a_i| is the condition number of the
    summation, used to state the floating-point tolerance of a reordered f32 sum."""
    affected, affecting, dt, d, sfx = _prep(affected, affecting)
    out = np.zeros(affected.shape[0], dtype=np.float64)
    getattr(lib(), f"oracle_bruteforce_abs_{sfx}")(
        _ptr(affected), C.c_size_t(len(affected)), _ptr(affecting), C.c_size_t(len(affecting)),
        _scalar(dt, softening), _ptr(out))
    return out


def brute_force_simd8_parallel(affected, affecting, softening=0.0, checked=True):
    """parallel::BruteForceSimd<8> restated with AVX2 + OpenMP (timing baseline, f32 3-D only)."""
    affected, affecting, dt, d, sfx = _prep(affected, affecting)
    assert sfx == "f32x3"
    out = np.zeros((affected.shape[0], d), dtype=dt)
    rc = lib().oracle_bruteforce_simd8_parallel_f32x3(
        _ptr(affected), C.c_size_t(len(affected)), _ptr(affecting), C.c_size_t(len(affecting)),
        C.c_float(softening), C.c_int(int(checked)), _ptr(out))
    assert rc == 0
    return out


def use_all_cores() -> int:
    """Use every host core for the OpenMP legs (torchrun sets OMP_NUM_THREADS=1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().oracle_set_threads(C.c_int(n))
    return n


def baseline_threads() -> int:
    return int(lib().oracle_baseline_threads())


class Tree:
    """RootedOrthtree restated (storage.rs:11-46, tree/mod.rs:91-138)."""

    def __init__(self, affecting):
        affecting = np.ascontiguousarray(affecting)
        self.dt = affecting.dtype
        self.d = affecting.shape[1] - 1
        self.sfx = _SFX[(self.dt, self.d)]
        fn = getattr(lib(), f"oracle_tree_build_{self.sfx}")
        fn.restype = C.c_void_p
        self.h = C.c_void_p(fn(_ptr(affecting), C.c_size_t(len(affecting))))
        info = np.zeros(3, dtype=np.int64)
        getattr(lib(), f"oracle_tree_info_{self.sfx}")(self.h, _ptr(info))
        self.n_nodes, self.max_depth, self.depth_limit_hit = (int(v) for v in info)

    def data(self):
        out = np.zeros((self.n_nodes, self.d + 1), dtype=self.dt)
        getattr(lib(), f"oracle_tree_data_{self.sfx}")(self.h, _ptr(out))
        return out

    def traverse(self, affected, theta, softening=0.0, checked=True, parallel=False,
                 counters=None):
        affected = np.ascontiguousarray(affected, dtype=self.dt)
        out = np.zeros((affected.shape[0], self.d), dtype=self.dt)
        if parallel:
            getattr(lib(), f"oracle_tree_traverse_parallel_{self.sfx}")(
                self.h, _ptr(affected), C.c_size_t(len(affected)), _scalar(self.dt, theta),
                _scalar(self.dt, softening), C.c_int(int(checked)), _ptr(out))
        else:
            cnt = np.zeros(2, dtype=np.int64)
            getattr(lib(), f"oracle_tree_traverse_{self.sfx}")(
                self.h, _ptr(affected), C.c_size_t(len(affected)), _scalar(self.dt, theta),
                _scalar(self.dt, softening), C.c_int(int(checked)), _ptr(out), _ptr(cnt))
            if counters is not None:
                counters["visits"] = int(cnt[0])
                counters["interactions"] = int(cnt[1])
        return out

    def __del__(self):
        try:
            getattr(lib(), f"oracle_tree_free_{self.sfx}")(self.h)
        except Exception:
            pass


def barnes_hut(affected, affecting, theta, softening=0.0, checked=True, parallel=False,
               counters=None):
    """sequential::BarnesHut / parallel::BarnesHut over Between(affected, affecting)
    (sequential.rs:526-543, parallel.rs:342-367): build, then traverse per affected."""
    return Tree(affecting).traverse(affected, theta, softening, checked, parallel, counters)


# ---- storage semantics (storage.rs:61-95, 153-163, 207-241) -----------------------------------
def between_of_slice(particles):
    """&[P] => Between(slice, slice) (storage.rs:231-241)."""
    p = np.ascontiguousarray(particles)
    return p[:, :-1], p


def between_of_reordered(particles):
    """&Reordered(particles, |p| mu != 0) => Between(unordered (original order), affecting copy)
    (storage.rs:153-163, 219-229; predicate GravitationalField::is_affecting, gravity/mod.rs:29-34)."""
    p = np.ascontiguousarray(particles)
    return p[:, :-1], np.ascontiguousarray(p[p[:, -1] != 0])


def between_of_ordered(particles):
    """&Ordered::new(particles, mu != 0) => Between(ordered_all, affecting prefix)
    (storage.rs:61-95, 207-217): stable partition, affecting first."""
    p = np.ascontiguousarray(particles)
    mask = p[:, -1] != 0
    ordered = np.concatenate([p[mask], p[~mask]])
    k = int(mask.sum())
    return ordered[:, :-1], np.ascontiguousarray(ordered[:k])


# ---- our tree specification (Morton keys, sort, linear orthtree) ------------------------------
NLEAF_DEFAULT = 16


class Octree:
    """CPU statement of the library's own tree specification (parity unpinned by the reference)."""

    def __init__(self, affecting, nleaf=NLEAF_DEFAULT, frame=None):
        """frame = (origin[D], ext, inv): build over `affecting` in the quantisation frame of a
        larger cloud (one part of the key-range-partitioned multi-GPU build)."""
        affecting = np.ascontiguousarray(affecting, dtype=np.float32)
        self.d = affecting.shape[1] - 1
        self.sfx = _SFX[(np.dtype(np.float32), self.d)]
        if frame is None:
            fn = getattr(lib(), f"oracle_octree_build_{self.sfx}")
            fn.restype = C.c_void_p
            self.h = C.c_void_p(fn(_ptr(affecting), C.c_size_t(len(affecting)), C.c_uint32(nleaf)))
        else:
            fr = np.ascontiguousarray(np.concatenate([np.asarray(frame[0], np.float32).ravel(),
                                                      np.array([frame[1], frame[2]], np.float32)]))
            fn = getattr(lib(), f"oracle_octree_build_in_frame_{self.sfx}")
            fn.restype = C.c_void_p
            self.h = C.c_void_p(fn(_ptr(affecting), C.c_size_t(len(affecting)), C.c_uint32(nleaf), _ptr(fr)))
        info = np.zeros(3, dtype=np.int64)
        frame = np.zeros(self.d + 2, dtype=np.float32)
        getattr(lib(), f"oracle_octree_info_{self.sfx}")(self.h, _ptr(info), _ptr(frame))
        self.n_nodes, self.n_levels, self.n = (int(v) for v in info)
        self.origin = frame[: self.d].copy()
        self.ext = float(frame[self.d])
        self.inv = float(frame[self.d + 1])
        n, m = self.n, self.n_nodes
        self.keys = np.zeros(n, dtype=np.uint64)
        self.perm = np.zeros(n, dtype=np.uint32)
        self.begin = np.zeros(m, dtype=np.uint32)
        self.count = np.zeros(m, dtype=np.uint32)
        self.level = np.zeros(m, dtype=np.uint32)
        self.first_child = np.zeros(m, dtype=np.uint32)
        self.n_child = np.zeros(m, dtype=np.uint32)
        self.commass = np.zeros((m, self.d + 1), dtype=np.float32)
        getattr(lib(), f"oracle_octree_read_{self.sfx}")(
            self.h, _ptr(self.keys), _ptr(self.perm), _ptr(self.begin), _ptr(self.count),
            _ptr(self.level), _ptr(self.first_child), _ptr(self.n_child), _ptr(self.commass))

    def moments(self):
        """Double-precision {sum m x_k, sum m} per node."""
        mom = np.zeros((self.n_nodes, self.d + 1), dtype=np.float64)
        getattr(lib(), f"oracle_octree_read_moments_{self.sfx}")(self.h, _ptr(mom))
        return mom

    def __del__(self):
        try:
            getattr(lib(), f"oracle_octree_free_{self.sfx}")(self.h)
        except Exception:
            pass


def morton_keys(positions, origin, inv):
    """Keys of bare positions (n, D) in a given quantisation frame."""
    positions = np.ascontiguousarray(positions, dtype=np.float32)
    d = positions.shape[1]
    sfx = _SFX[(np.dtype(np.float32), d)]
    keys = np.zeros(len(positions), dtype=np.uint64)
    origin = np.ascontiguousarray(origin, dtype=np.float32)
    getattr(lib(), f"oracle_morton_keys_{sfx}")(
        _ptr(positions), C.c_size_t(len(positions)), C.c_size_t(d), _ptr(origin), C.c_float(inv),
        _ptr(keys))
    return keys


# ---- the loop every caller writes around compute() -------------------------------------------------
def semi_implicit_euler(compute, particles, velocities, dt, steps, massive_only=False):
    """``steps`` times: a = compute(affected positions, affecting particles);
    velocity += a * dt; position += velocity * dt — examples/simple/src/main.rs:45-59 and the
    reference's circular_orbit! test (gravity/newtonian/mod.rs:318-331), each operation rounded
    separately in the particle scalar type (Rust never fuses a * b + c).

    ``compute(affected, affecting)`` is one of the oracle algorithms above.  massive_only: the
    `Reordered` storage — all particles affected, those with mu != 0 affecting
    (storage.rs:153-163, 219-229).  Returns (particles, velocities, last accelerations)."""
    p = np.array(particles, copy=True)
    dt_s = p.dtype.type(dt)
    d = p.shape[1] - 1
    v = np.zeros((len(p), d), p.dtype) if velocities is None else np.array(velocities, dtype=p.dtype)
    a = np.zeros((len(p), d), p.dtype)
    for _ in range(steps):
        src = np.ascontiguousarray(p[p[:, -1] != 0]) if massive_only else p
        a = compute(np.ascontiguousarray(p[:, :d]), src)
        v = v + a * dt_s          # numpy keeps the dtype: two roundings, like the Rust expression
        p[:, :d] = p[:, :d] + v * dt_s
    return p, v, a
