"""ctypes binding of libfcvm.so (include/fcvm.h).  The library is built in-tree by build.py;
loading fails loudly when it is missing — there is no Python or CPU fallback for any stage."""
import ctypes
import os

import numpy as np

from .params import Camera, Params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i8p = ctypes.POINTER(ctypes.c_int8)
i32p = ctypes.POINTER(ctypes.c_int32)
i64p = ctypes.POINTER(ctypes.c_int64)
u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
f32p = ctypes.POINTER(ctypes.c_float)
f64p = ctypes.POINTER(ctypes.c_double)
vp = ctypes.c_void_p
i32, i64, f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double

# fcvm_comm_fn: (user, op, buf_dev, n, stream) -> int
COMM_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)
COMM_ALLGATHER, COMM_ALLREDUCE_MAX = 0, 1

# name -> (restype, argtypes): every symbol include/fcvm.h declares
SYMBOLS = {
    "fcvm_abi_version": (ctypes.c_int, []),
    "fcvm_last_error": (ctypes.c_char_p, []),
    "fcvm_create": (vp, [ctypes.POINTER(Params)]),
    "fcvm_destroy": (None, [vp]),
    "fcvm_reset": (ctypes.c_int, [vp]),
    "fcvm_set_map_cloud": (ctypes.c_int, [vp, f32p, i64]),
    "fcvm_set_model": (ctypes.c_int, [vp, f32p, i64]),
    "fcvm_set_normals": (ctypes.c_int, [vp, f64p, f64p, i64]),
    "fcvm_set_init_viewpoints": (ctypes.c_int, [vp, f32p, i64]),
    "fcvm_update": (ctypes.c_int, [vp]),
    "fcvm_get_viewpoints": (i64, [vp, i32p, f64p, i32p, i64]),
    "fcvm_get_uncovered": (i64, [vp, f32p, i64]),
    "fcvm_get_vps_pose": (i64, [vp, f64p, i64]),
    "fcvm_get_vps_voxcount": (i64, [vp, i32p, i64]),
    "fcvm_get_vps_contri": (i64, [vp, i32p, i64]),
    "fcvm_get_cover_state": (i64, [vp, u8p, i32p, i32p, i64]),
    "fcvm_get_grid": (ctypes.c_int, [vp, f64p, i32p, u8p, i64]),
    "fcvm_safety_check": (ctypes.c_int, [vp, f32p, i64, u8p]),
    "fcvm_trace_enable": (ctypes.c_int, [vp, i32]),
    "fcvm_trace_count": (i64, [vp]),
    "fcvm_trace_words": (i64, [vp]),
    "fcvm_trace_get": (ctypes.c_int, [vp, i64, i32p, i32p, f64p, u32p]),
    "fcvm_get_counters": (ctypes.c_int, [vp, i64p]),
    "fcvm_evaluate": (ctypes.c_int, [vp, i32, f64p, i64, u32p, i32p, u32p]),
    "fcvm_evaluate_csr": (i64, [vp, i32, f64p, i64, u32p, i32p, i64p, u32p, i64]),
    "fcvm_stage_poses": (ctypes.c_int, [vp, f64p, i64]),
    "fcvm_evaluate_resident": (ctypes.c_int, [vp, i32, vp, i32, i64p]),
    "fcvm_resident_buffers": (ctypes.c_int, [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), i64p, i64p]),
    "fcvm_own_resident": (ctypes.c_int, [vp, i32p, i32p, i64]),
    "fcvm_host_alloc": (vp, [i64]),
    "fcvm_host_free": (None, [vp]),
    "fcvm_last_kernel_ms": (f64, [vp]),
    "fcvm_debug_bounds_violations": (i64, [vp]),
    "fcvm_set_shard": (ctypes.c_int, [vp, i32, i32, COMM_FN, vp]),
    "fcvm_nccl_unique_id": (ctypes.c_int, [u8p]),
    "fcvm_set_shard_nccl": (ctypes.c_int, [vp, u8p, i32, i32]),
    "fcvm_comm_stats": (ctypes.c_int, [vp, i64p]),
    "fcvm_get_timers": (ctypes.c_int, [vp, f64p]),
    "fcvm_host_fov_normals": (ctypes.c_int, [ctypes.POINTER(Params), f64, f64, f64p]),
    "fcvm_host_build_grid": (ctypes.c_int, [ctypes.POINTER(Params), f32p, i64, f64p, i32p, u8p, i64]),
    "fcvm_shard_range": (ctypes.c_int, [i64, i32, i32, i64p, i64p]),
    "fcvm_grid_create": (vp, [i32, f64p, f64, i32p, i8p, i8p, i8p]),
    "fcvm_grid_destroy": (None, [vp]),
    "fcvm_grid_light_state": (ctypes.c_int, [vp, f64p, i64, f32p, i64, u8p, i32p, i32p]),
    "fcvm_grid_safety_box": (ctypes.c_int, [vp, f32p, i64, f64, u8p]),
    "fcvm_grid_line_of_sight": (ctypes.c_int, [vp, f64p, i64, u8p, f64p]),
    "fcvm_grid_set_occ": (ctypes.c_int, [vp, f32p, i64]),
    "fcvm_grid_internal_space": (ctypes.c_int, [vp, f32p, i64p, i64, f64p, i32, f64, f64]),
    "fcvm_grid_outer_check": (ctypes.c_int, [vp, f64p, i64, f64]),
    "fcvm_grid_get_buffers": (ctypes.c_int, [vp, i8p, i8p, i8p]),
    "fcvm_grid_oob_writes": (i64, [vp]),
    "fcvm_grid_counters": (ctypes.c_int, [vp, i64p]),
    "fcvm_grid_last_kernel_ms": (f64, [vp]),
    "fcvm_cov_create": (vp, [ctypes.POINTER(Params), ctypes.POINTER(Camera)]),
    "fcvm_cov_destroy": (None, [vp]),
    "fcvm_cov_set_cloud": (ctypes.c_int, [vp, f32p, i64]),
    "fcvm_cov_pinhole": (i64, [vp, f64p, i64, i64p, i32p, i64, u8p]),
    "fcvm_cov_evaluate": (i64, [vp, f64p, i64, f64p, i64p, i32p, i64, u8p]),
    "fcvm_cov_fetch_ids": (i64, [vp, i32p, i64]),
    "fcvm_cov_counters": (ctypes.c_int, [vp, i64p]),
    "fcvm_cov_chunk_points": (ctypes.c_int32, []),
    "fcvm_cov_last_kernel_ms": (f64, [vp]),
    "fcvm_cov_last_scan_ms": (f64, [vp]),
}


def lib_path():
    return os.environ.get("FCVM_LIB") or os.path.join(_HERE, "_build", "libfcvm.so")


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python fc-planner_b200/build.py` (or __graft_entry__.build()). "
                           "The CUDA library is the only implementation; there is no fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError = the library does not export what fcvm.h declares
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


class FcvmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fcvm error {code}: {msg}")
        self.code = code


def check(rc):
    if rc < 0:
        raise FcvmError(rc, load().fcvm_last_error().decode("utf-8", "replace"))
    return rc
