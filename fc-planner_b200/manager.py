"""Host-side mirror of `predrecon::ViewpointManager` over the C ABI (include/fcvm.h).

Same method names, argument meaning and error behaviour as the reference's public interface
(viewpoint_manager/include/viewpoint_manager/viewpoint_manager.h:180-193): clouds are numpy arrays
instead of pcl::PointCloud, and, like the reference, bad input is logged and the call returns
without raising (`last_status` keeps the code); a failing CUDA path raises — it never degrades to
a CPU implementation.
"""
import ctypes
import logging

import numpy as np

from . import _lib
from ._lib import check, f32p, f64p, i32p, i64p, ptr, u8p, u32p

log = logging.getLogger("fcvm")

STAGE_E1, STAGE_E2, STAGE_E3, STAGE_E4 = 1, 2, 3, 4
_SOFT = (-1, -2)   # FCVM_EINVAL, FCVM_ENOTREADY: the reference logs ROS_ERROR and returns


class SingleViewpoint:
    """viewpoint_manager.h:79-84"""
    __slots__ = ("vp_id", "pose", "vox_count")

    def __init__(self, vp_id, pose, vox_count):
        self.vp_id, self.pose, self.vox_count = int(vp_id), np.array(pose, dtype=np.float64), int(vox_count)


class ViewpointManager:
    def __init__(self, params):
        """ViewpointManager() + init(nh): `params` stands for the ROS parameter server."""
        self.lib = _lib.load()
        self.params = params
        self.h = ctypes.c_void_p(self.lib.fcvm_create(ctypes.byref(params)))
        if not self.h:
            raise _lib.FcvmError(-1, self.lib.fcvm_last_error().decode())
        self.last_status = 0
        self.P = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.fcvm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _soft(self, rc):
        self.last_status = rc
        if rc in _SOFT:
            log.error("[ViewpointManager] %s", self.lib.fcvm_last_error().decode("utf-8", "replace"))
            return rc
        return check(rc)

    # ---- the reference's public interface ---------------------------------------------------
    def reset(self):
        return self._soft(self.lib.fcvm_reset(self.h))

    def setMapPointCloud(self, input_map):
        xyz = np.ascontiguousarray(input_map, dtype=np.float32).reshape(-1, 3)
        return self._soft(self.lib.fcvm_set_map_cloud(self.h, ptr(xyz, f32p), len(xyz)))

    def setModel(self, input_model):
        xyz = np.ascontiguousarray(input_model, dtype=np.float32).reshape(-1, 3)
        rc = self._soft(self.lib.fcvm_set_model(self.h, ptr(xyz, f32p), len(xyz)))
        if rc == 0:
            self.P = len(xyz)
        return rc

    def setNormals(self, points, normals):
        """pt_normal_pairs: keys = exact (float -> double) point coordinates, values = normals."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, 3)
        return self._soft(self.lib.fcvm_set_normals(self.h, ptr(pts, f64p), ptr(nrm, f64p), len(pts)))

    def setInitViewpoints(self, input_normal_vps):
        """rows of (x, y, z, normal_x, normal_y, normal_z) float32, like pcl::PointNormal."""
        v = np.ascontiguousarray(input_normal_vps, dtype=np.float32).reshape(-1, 6)
        return self._soft(self.lib.fcvm_set_init_viewpoints(self.h, ptr(v, f32p), len(v)))

    def updateViewpoints(self):
        return self._soft(self.lib.fcvm_update(self.h))

    def getUpdatedViewpoints(self):
        n = check(self.lib.fcvm_get_viewpoints(self.h, None, None, None, 0))
        ids = np.zeros(n, np.int32); pose = np.zeros((n, 5), np.float64); vox = np.zeros(n, np.int32)
        check(self.lib.fcvm_get_viewpoints(self.h, ptr(ids, i32p), ptr(pose, f64p), ptr(vox, i32p), n))
        return [SingleViewpoint(ids[i], pose[i], vox[i]) for i in range(n)]

    def getUncoveredModelCloud(self):
        n = check(self.lib.fcvm_get_uncovered(self.h, None, 0))
        out = np.zeros((n, 3), np.float32)
        check(self.lib.fcvm_get_uncovered(self.h, ptr(out, f32p), n))
        return out

    # ---- arrays-out variants and parity taps ------------------------------------------------
    def final_arrays(self):
        n = check(self.lib.fcvm_get_viewpoints(self.h, None, None, None, 0))
        ids = np.zeros(n, np.int32); pose = np.zeros((n, 5), np.float64); vox = np.zeros(n, np.int32)
        check(self.lib.fcvm_get_viewpoints(self.h, ptr(ids, i32p), ptr(pose, f64p), ptr(vox, i32p), n))
        return ids, pose, vox

    def vps_pose(self):
        n = check(self.lib.fcvm_get_vps_pose(self.h, None, 0))
        out = np.zeros((n, 5), np.float64)
        check(self.lib.fcvm_get_vps_pose(self.h, ptr(out, f64p), n))
        return out

    def vps_voxcount(self):
        n = check(self.lib.fcvm_get_vps_voxcount(self.h, None, 0))
        out = np.zeros(n, np.int32)
        check(self.lib.fcvm_get_vps_voxcount(self.h, ptr(out, i32p), n))
        return out

    def vps_contri(self):
        n = check(self.lib.fcvm_get_vps_contri(self.h, None, 0))
        out = np.zeros(n, np.int32)
        check(self.lib.fcvm_get_vps_contri(self.h, ptr(out, i32p), n))
        return out

    def cover_state(self):
        n = check(self.lib.fcvm_get_cover_state(self.h, None, None, None, 0))
        cov = np.zeros(n, np.uint8); con = np.zeros(n, np.int32); cid = np.zeros(n, np.int32)
        check(self.lib.fcvm_get_cover_state(self.h, ptr(cov, u8p), ptr(con, i32p), ptr(cid, i32p), n))
        return cov, con, cid

    def grid(self, with_bits=True):
        origin = np.zeros(3, np.float64); dims = np.zeros(3, np.int32)
        check(self.lib.fcvm_get_grid(self.h, ptr(origin, f64p), ptr(dims, i32p), None, 0))
        bits = None
        if with_bits:
            g = int(dims[0]) * int(dims[1]) * int(dims[2])
            bits = np.zeros((g + 7) // 8, np.uint8)
            check(self.lib.fcvm_get_grid(self.h, None, None, ptr(bits, u8p), len(bits)))
        return origin, dims, bits

    def trace_enable(self, on=True):
        check(self.lib.fcvm_trace_enable(self.h, 1 if on else 0))

    def words(self):
        return int(self.lib.fcvm_trace_words(self.h))

    def trace(self):
        n = self.lib.fcvm_trace_count(self.h)
        words = self.words()
        out = []
        for i in range(n):
            st = ctypes.c_int32(); vid = ctypes.c_int32()
            pose = np.zeros(5, np.float64); row = np.zeros(words, np.uint32)
            check(self.lib.fcvm_trace_get(self.h, i, ctypes.byref(st), ctypes.byref(vid), ptr(pose, f64p), ptr(row, u32p)))
            out.append((st.value, vid.value, pose, row))
        return out

    def counters(self):
        out = np.zeros(8, np.int64)
        check(self.lib.fcvm_get_counters(self.h, ptr(out, i64p)))
        return out

    def safety_check(self, xyz):
        """viewpointSafetyCheck (viewpoint_manager.cpp:889-946) of n candidate positions -> uint8[n]."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        ok = np.zeros(len(xyz), np.uint8)
        check(self.lib.fcvm_safety_check(self.h, ptr(xyz, f32p), len(xyz), ptr(ok, u8p)))
        return ok

    # ---- evaluator level ----------------------------------------------------------------------
    def evaluate(self, stage, pose5, restrict_rows=None, want_rows=True, out=None):
        """One evaluator pass over host poses -> (count[n], rows[n, words]); H2D + kernel + D2H.
        out = (count, rows) preallocated arrays (e.g. from pinned_empty) to receive the result."""
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        n, words = len(pose5), self.words()
        if out is not None:
            count, rows = out
            assert count.shape == (n,) and count.dtype == np.int32 and rows.shape == (n, words) and rows.dtype == np.uint32
        else:
            count = np.zeros(n, np.int32)
            rows = np.zeros((n, words), np.uint32) if want_rows else None
        if restrict_rows is not None:
            restrict_rows = np.ascontiguousarray(restrict_rows, dtype=np.uint32)
            assert restrict_rows.shape == (n, words)
        check(self.lib.fcvm_evaluate(self.h, stage, ptr(pose5, f64p), n, ptr(restrict_rows, u32p), ptr(count, i32p), ptr(rows, u32p)))
        return count, rows

    def evaluate_csr(self, stage, pose5, restrict_rows=None, cap=None, out=None):
        """Evaluator pass returning (count[n], offsets[n+1], indices[total]): the visible sets as CSR.
        out = (count, offsets, indices) preallocated arrays (e.g. from pinned_empty: the index lists are then copied
        straight from the device into `indices` while the next sub-batch is being evaluated)."""
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        n = len(pose5)
        if restrict_rows is not None:
            restrict_rows = np.ascontiguousarray(restrict_rows, dtype=np.uint32)
        if out is not None:
            count, offsets, indices = out
            total = check(self.lib.fcvm_evaluate_csr(self.h, stage, ptr(pose5, f64p), n, ptr(restrict_rows, u32p), ptr(count, i32p),
                                                     ptr(offsets, i64p), ptr(indices, u32p), len(indices)))
            if total > len(indices):
                raise ValueError(f"indices buffer too small: {total} visible pairs, room for {len(indices)}")
            return count, offsets, indices[:total]
        count = np.zeros(n, np.int32); offsets = np.zeros(n + 1, np.int64)
        cap = int(cap) if cap else max(1, 64 * n)
        while True:
            indices = np.empty(cap, np.uint32)
            total = check(self.lib.fcvm_evaluate_csr(self.h, stage, ptr(pose5, f64p), n, ptr(restrict_rows, u32p), ptr(count, i32p),
                                                     ptr(offsets, i64p), ptr(indices, u32p), cap))
            if total <= cap:
                return count, offsets, indices[:total]
            cap = int(total)

    def stage_poses(self, pose5):
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        check(self.lib.fcvm_stage_poses(self.h, ptr(pose5, f64p), len(pose5)))

    def evaluate_resident(self, stage, stream=None, sync=True):
        rays = ctypes.c_int64(0)
        check(self.lib.fcvm_evaluate_resident(self.h, stage, stream, 1 if sync else 0, ctypes.byref(rays)))
        return rays.value

    def resident_buffers(self):
        rows = ctypes.c_void_p(); cnt = ctypes.c_void_p(); n = ctypes.c_int64(); w = ctypes.c_int64()
        check(self.lib.fcvm_resident_buffers(self.h, ctypes.byref(rows), ctypes.byref(cnt), ctypes.byref(n), ctypes.byref(w)))
        return rows.value, cnt.value, n.value, w.value

    def own_resident(self, vp_ids, order=None):
        vp_ids = np.ascontiguousarray(vp_ids, dtype=np.int32)
        order = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
        check(self.lib.fcvm_own_resident(self.h, ptr(vp_ids, i32p), ptr(order, i32p), len(vp_ids)))

    def debug_bounds_violations(self):
        """-1 unless the loaded library is the FCVM_DEBUG_BOUNDS build (libfcvm_dbg.so)."""
        return int(self.lib.fcvm_debug_bounds_violations(self.h))

    def last_kernel_ms(self):
        return float(self.lib.fcvm_last_kernel_ms(self.h))

    def set_shard(self, rank, world, comm_cb=None):
        """Caller-supplied collectives: comm_cb(user, op, buf_dev, n, stream) -> 0 (include/fcvm.h: fcvm_comm_fn)."""
        cb = _lib.COMM_FN(comm_cb) if comm_cb is not None else ctypes.cast(None, _lib.COMM_FN)
        self._keep.append(cb)
        check(self.lib.fcvm_set_shard(self.h, rank, world, cb, None))

    def set_shard_nccl(self, unique_id, rank, world):
        """The library drives NCCL itself; unique_id = the 128 bytes fcvm_nccl_unique_id() produced on one rank."""
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert uid.size == 128
        check(self.lib.fcvm_set_shard_nccl(self.h, ptr(uid, u8p), rank, world))

    def comm_stats(self):
        """(all-gathers, bytes gathered, all-reduces, bytes reduced) issued by this handle since create / reset."""
        out = np.zeros(4, np.int64)
        check(self.lib.fcvm_comm_stats(self.h, ptr(out, i64p)))
        return out

    TIMER_NAMES = ("setInitViewpoints", "updateViewpoints", "evaluator_batches", "ownership", "uncovered_candidates", "host_libm",
                   "safety_batches", "pose_sum_pipeline", "collectives_enqueue", "k_eval_E1", "k_eval_E2", "k_eval_E3", "k_eval_E4")

    def timers(self):
        """Wall-clock split of the manager-level calls since create / reset, ms (fcvm_get_timers)."""
        out = np.zeros(13, np.float64)
        check(self.lib.fcvm_get_timers(self.h, ptr(out, f64p)))
        return dict(zip(self.TIMER_NAMES, (float(x) for x in out)))


class _Pinned:
    def __init__(self, nbytes):
        self.lib = _lib.load()
        self.ptr = self.lib.fcvm_host_alloc(nbytes)
        if not self.ptr:
            raise _lib.FcvmError(-4, self.lib.fcvm_last_error().decode("utf-8", "replace"))

    def __del__(self):
        try:
            self.lib.fcvm_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory (fcvm_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    owner = _Pinned(max(1, n * dtype.itemsize))
    buf = (ctypes.c_uint8 * (n * dtype.itemsize)).from_address(owner.ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _PINNED_OWNERS[id(buf)] = owner
    import weakref
    weakref.finalize(buf, _PINNED_OWNERS.pop, id(buf), None)
    return arr


_PINNED_OWNERS = {}


def host_fov_normals(params, pitch, yaw):
    out = np.zeros(12, np.float64)
    check(_lib.load().fcvm_host_fov_normals(ctypes.byref(params), pitch, yaw, ptr(out, f64p)))
    return out.reshape(4, 3)


def host_build_grid(params, xyz, with_bits=True):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    origin = np.zeros(3, np.float64); dims = np.zeros(3, np.int32)
    check(_lib.load().fcvm_host_build_grid(ctypes.byref(params), ptr(xyz, f32p), len(xyz), ptr(origin, f64p), ptr(dims, i32p), None, 0))
    bits = None
    if with_bits:
        g = int(dims[0]) * int(dims[1]) * int(dims[2])
        bits = np.zeros((g + 7) // 8, np.uint8)
        check(_lib.load().fcvm_host_build_grid(ctypes.byref(params), ptr(xyz, f32p), len(xyz), None, None, ptr(bits, u8p), len(bits)))
    return origin, dims, bits


def nccl_unique_id():
    """128 bytes of a fresh ncclUniqueId (fcvm_nccl_unique_id): make it on one rank, hand it to the others."""
    out = np.zeros(128, np.uint8)
    check(_lib.load().fcvm_nccl_unique_id(ptr(out, u8p)))
    return out.tobytes()


def shard_range(n, rank, world):
    lo = ctypes.c_int64(); hi = ctypes.c_int64()
    check(_lib.load().fcvm_shard_range(n, rank, world, ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value
