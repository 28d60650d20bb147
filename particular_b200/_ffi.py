"""ctypes binding of libparticular_cuda.so (C ABI: include/particular_cuda.h).

There is no CPU fallback: importing this module without the built library raises, and creating a
context without a B200-class device raises.  Build with ``python -m particular_b200.build`` or
``__graft_entry__.build()``.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libparticular_cuda.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA backend is not built and there is no CPU fallback. "
        "Run `python -m particular_b200.build`.")

lib = C.CDLL(LIB_PATH)

ABI_VERSION = 1
FLAG_NO_PHASE_TIMINGS = 1
FLAG_BH_PARTITIONED_BUILD = 2
FLAG_BH_REPLICATED_BUILD = 4
FLAG_EXACT_CHECKED = 8
FLAG_BH_LET_BUILD = 16
UNIQUE_ID_BYTES = 128

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_NO_DEVICE = -2
ERR_CUDA = -3
ERR_OUT_OF_MEMORY = -4
ERR_NCCL = -5
ERR_TREE_OVERFLOW = -6
ERR_NOT_INITIALISED = -7


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("leaf_size", C.c_uint32),
                ("expansion_order", C.c_uint32)]


class Timings(C.Structure):
    _fields_ = [("upload_ms", C.c_float), ("comm_ms", C.c_float), ("build_ms", C.c_float),
                ("compute_ms", C.c_float), ("download_ms", C.c_float),
                ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class TreeInfo(C.Structure):
    _fields_ = [("n_particles", C.c_uint64), ("n_nodes", C.c_uint64), ("n_levels", C.c_uint32),
                ("leaf_size", C.c_uint32), ("dim", C.c_uint32), ("bits", C.c_uint32),
                ("origin", C.c_float * 3), ("extent", C.c_float), ("inv", C.c_float),
                ("reserved", C.c_uint32)]


class SimConfig(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("scalar", C.c_uint32), ("algorithm", C.c_uint32),
                ("flags", C.c_uint32), ("theta", C.c_double), ("softening", C.c_double),
                ("dt", C.c_double), ("checked", C.c_int32), ("reserved", C.c_uint32)]


class SimInfo(C.Structure):
    _fields_ = [("n_particles", C.c_uint64), ("n_affecting", C.c_uint64), ("steps_done", C.c_uint64),
                ("d_particles", C.c_void_p), ("d_velocities", C.c_void_p),
                ("d_accelerations", C.c_void_p), ("graph_active", C.c_uint32),
                ("launches_per_step", C.c_uint32)]


BRUTE_FORCE, BARNES_HUT = 0, 1
F32, F64 = 0, 1
SIM_AFFECTING_MASSIVE_ONLY, SIM_NO_GRAPH = 1, 2

TREE_KEYS, TREE_PERM, TREE_NODE_BEGIN, TREE_NODE_COUNT, TREE_NODE_LEVEL, TREE_NODE_FIRST_CHILD, \
    TREE_NODE_NUM_CHILDREN, TREE_NODE_COM_MASS = range(8)

_vp, _sz, _f, _d, _i = C.c_void_p, C.c_size_t, C.c_float, C.c_double, C.c_int

# name -> (restype, argtypes); every symbol include/particular_cuda.h declares.
SIGNATURES = {
    "pcuda_abi_version": (_i, []),
    "pcuda_status_string": (C.c_char_p, [_i]),
    "pcuda_device_count": (_i, [C.POINTER(_i)]),
    "pcuda_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "pcuda_destroy": (None, [_vp]),
    "pcuda_last_error": (C.c_char_p, [_vp]),
    "pcuda_get_timings": (_i, [_vp, C.POINTER(Timings)]),
    "pcuda_stream": (_vp, [_vp]),
    "pcuda_sync": (_i, [_vp]),
    "pcuda_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.c_char_p, _sz]),
    "pcuda_host_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "pcuda_host_free": (_i, [_vp, _vp]),
    "pcuda_bruteforce_f32x3": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f32x2": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f64x3": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _i, _vp]),
    "pcuda_bruteforce_f32x3_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f32x2_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f64x3_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _i, _vp]),
    "pcuda_bruteforce_f64x2": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _i, _vp]),
    "pcuda_bruteforce_f64x2_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _i, _vp]),
    "pcuda_barneshut_f32x3": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _f, _i, _vp]),
    "pcuda_barneshut_f32x2": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _f, _i, _vp]),
    "pcuda_barneshut_f32x3_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _f, _i, _vp]),
    "pcuda_barneshut_f32x2_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _f, _f, _i, _vp]),
    "pcuda_barneshut_f64x3": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _d, _i, _vp]),
    "pcuda_barneshut_f64x2": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _d, _i, _vp]),
    "pcuda_barneshut_f64x3_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _d, _i, _vp]),
    "pcuda_barneshut_f64x2_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _d, _d, _i, _vp]),
    "pcuda_tree_build_f32": (_i, [_vp, C.c_uint32, _vp, _sz, C.POINTER(_vp)]),
    "pcuda_tree_info_get": (_i, [_vp, C.POINTER(TreeInfo)]),
    "pcuda_tree_read": (_i, [_vp, _vp, _i, _vp, _sz]),
    "pcuda_tree_traverse_f32": (_i, [_vp, _vp, _vp, _sz, _f, _f, _i, _vp]),
    "pcuda_tree_last_counters": (_i, [_vp, C.POINTER(C.c_uint64 * 5)]),
    "pcuda_tree_destroy": (None, [_vp, _vp]),
    "pcuda_comm_unique_id": (_i, [_vp, C.POINTER(C.c_uint8 * UNIQUE_ID_BYTES)]),
    "pcuda_comm_init": (_i, [_vp, C.POINTER(C.c_uint8 * UNIQUE_ID_BYTES), _i, _i]),
    "pcuda_comm_destroy": (_i, [_vp]),
    "pcuda_comm_init_local": (_i, [C.POINTER(_vp), _i]),
    "pcuda_comm_allgather_dev": (_i, [_vp, _vp, _vp, _sz]),
    "pcuda_bruteforce_f32x3_sharded": (_i, [_vp, _vp, _sz, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f32x3_sharded_dev": (_i, [_vp, _vp, _sz, _sz, _f, _i, _vp, _vp]),
    "pcuda_bruteforce_f32x3_between_sharded": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _f, _i, _vp]),
    "pcuda_bruteforce_f32x3_between_sharded_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _f, _i, _vp, _vp]),
    "pcuda_morton_f32x3": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "pcuda_morton_f32x2": (_i, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "pcuda_barneshut_f32x3_sharded_dev": (_i, [_vp, _vp, _sz, _sz, _f, _f, _i, _vp, _vp]),
    "pcuda_barneshut_f32x3_sharded": (_i, [_vp, _vp, _sz, _sz, _f, _f, _i, _vp]),
    "pcuda_barneshut_f32x3_partitioned": (_i, [_vp, _vp, _sz, _i, _f, _f, _i, _vp]),
    "pcuda_barneshut_f32x3_partitioned_dev": (_i, [_vp, _vp, _sz, _i, _f, _f, _i, _vp]),
    "pcuda_sim_create": (_i, [_vp, C.POINTER(SimConfig), _vp, _vp, _sz, C.POINTER(_vp)]),
    "pcuda_sim_configure": (_i, [_vp, _vp, C.POINTER(SimConfig)]),
    "pcuda_sim_step": (_i, [_vp, _vp, C.c_uint32]),
    "pcuda_sim_read": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "pcuda_sim_info": (_i, [_vp, C.POINTER(SimInfo)]),
    "pcuda_sim_destroy": (None, [_vp, _vp]),
    "pcuda_interaction_check": (_i, [C.c_char_p, C.c_char_p, _sz]),
    "pcuda_interaction_create": (_i, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "pcuda_interaction_sizes": (_i, [_vp, C.POINTER(C.c_uint32 * 4)]),
    "pcuda_interaction_brute_force": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp]),
    "pcuda_interaction_destroy": (None, [_vp, _vp]),
    # not in the stable header: measurement / tuning hooks
    "pcuda_probe_fp32": (_i, [_vp, _i, _i, _i, C.POINTER(_d), C.POINTER(_f)]),
    "pcuda_debug_set": (_i, [C.c_char_p, _i]),
    "pcuda_debug_merge_top_tree": (_i, [_i, _vp, _vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, C.c_uint32, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export the symbol
    _fn.restype = _res
    _fn.argtypes = _args

if lib.pcuda_abi_version() != ABI_VERSION:
    raise ImportError("libparticular_cuda.so ABI version mismatch")


class CudaError(RuntimeError):
    """A C-ABI call returned a negative status (the Rust wrapper would panic here, matching the
    reference's unwrap()/expect() behaviour, gpu/resources.rs:38-39, 341-345)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[{status}: {lib.pcuda_status_string(status).decode()}] {message}")
        self.status = status


def check(status: int, ctx=None) -> None:
    if status != OK:
        raise CudaError(status, (lib.pcuda_last_error(ctx) or b"").decode())
