"""ctypes front end of the CPU ORACLE (oracle/fcvm_oracle.hpp).  TEST INFRASTRUCTURE.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() import
this module, and only as the checker.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists) with the committed Makefile."""
    so = os.path.join(_HERE, "_build", "libfcvm_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("fcvm_oracle.hpp", "fcvm_oracle_cov.hpp", "fcvm_oracle_plan.hpp", "oracle_capi.cpp")):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
    return so


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def load(alt=False):
    name = "libfcvm_oracle_alt.so" if alt else "libfcvm_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    build()
    lib = ctypes.CDLL(os.path.join(_HERE, "_build", name))
    lib.orc_create.restype = ctypes.c_void_p
    lib.orc_create.argtypes = [ctypes.c_void_p]
    lib.orc_destroy.argtypes = [ctypes.c_void_p]
    for fn in ("orc_get_viewpoints", "orc_get_uncovered", "orc_get_vps_pose", "orc_get_vps_voxcount", "orc_get_vps_contri",
               "orc_get_cover_state", "orc_trace_count", "orc_trace_words", "orc_prune_order", "orc_ray_ids", "orc_biray_ids"):
        getattr(lib, fn).restype = ctypes.c_int64
    lib.orc_cov_create.restype = ctypes.c_void_p
    lib.orc_cov_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.orc_cov_destroy.argtypes = [ctypes.c_void_p]
    lib.orc_cov_pinhole.restype = ctypes.c_int64
    lib.orc_cov_evaluate.restype = ctypes.c_int64
    lib.orc_plan_create.restype = ctypes.c_void_p
    lib.orc_plan_destroy.argtypes = [ctypes.c_void_p]
    lib.orc_intbound.restype = ctypes.c_double
    lib.orc_intbound.argtypes = [ctypes.c_double, ctypes.c_double]
    lib.orc_mod.restype = ctypes.c_double
    lib.orc_mod.argtypes = [ctypes.c_double, ctypes.c_double]
    _LIBS[name] = lib
    return lib


class Oracle:
    """The reference's ViewpointManager call sequence on numpy arrays, run by the CPU oracle."""

    def __init__(self, params, alt=False):
        self.lib = load(alt)
        self.params = params
        self.h = ctypes.c_void_p(self.lib.orc_create(ctypes.byref(params)))
        self.P = 0

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the 8-call sequence ---------------------------------------------------------------
    def reset(self):
        self.lib.orc_reset(self.h)

    def setMapPointCloud(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        return self.lib.orc_set_map_cloud(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def setModel(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.P = len(xyz)
        return self.lib.orc_set_model(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def setNormals(self, pts, normals):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        return self.lib.orc_set_normals(self.h, _ptr(pts, c_f64p), _ptr(normals, c_f64p), ctypes.c_int64(len(pts)))

    def setInitViewpoints(self, pos_dir):
        pos_dir = np.ascontiguousarray(pos_dir, dtype=np.float32)
        return self.lib.orc_set_init_viewpoints(self.h, _ptr(pos_dir, c_f32p), ctypes.c_int64(len(pos_dir)))

    def updateViewpoints(self):
        return self.lib.orc_update(self.h)

    def getUpdatedViewpoints(self):
        n = self.lib.orc_get_viewpoints(self.h, None, None, None, ctypes.c_int64(0))
        ids = np.zeros(n, np.int32); pose = np.zeros((n, 5), np.float64); vox = np.zeros(n, np.int32)
        self.lib.orc_get_viewpoints(self.h, _ptr(ids, c_i32p), _ptr(pose, c_f64p), _ptr(vox, c_i32p), ctypes.c_int64(n))
        return ids, pose, vox

    def getUncoveredModelCloud(self):
        n = self.lib.orc_get_uncovered(self.h, None, ctypes.c_int64(0))
        out = np.zeros((n, 3), np.float32)
        self.lib.orc_get_uncovered(self.h, _ptr(out, c_f32p), ctypes.c_int64(n))
        return out

    # -- taps --------------------------------------------------------------------------------
    def vps_pose(self):
        n = self.lib.orc_get_vps_pose(self.h, None, ctypes.c_int64(0))
        out = np.zeros((n, 5), np.float64)
        self.lib.orc_get_vps_pose(self.h, _ptr(out, c_f64p), ctypes.c_int64(n))
        return out

    def vps_voxcount(self):
        n = self.lib.orc_get_vps_voxcount(self.h, None, ctypes.c_int64(0))
        out = np.zeros(n, np.int32)
        self.lib.orc_get_vps_voxcount(self.h, _ptr(out, c_i32p), ctypes.c_int64(n))
        return out

    def vps_contri(self):
        n = self.lib.orc_get_vps_contri(self.h, None, ctypes.c_int64(0))
        out = np.zeros(n, np.int32)
        self.lib.orc_get_vps_contri(self.h, _ptr(out, c_i32p), ctypes.c_int64(n))
        return out

    def cover_state(self):
        n = self.lib.orc_get_cover_state(self.h, None, None, None, ctypes.c_int64(0))
        cov = np.zeros(n, np.uint8); con = np.zeros(n, np.int32); cid = np.zeros(n, np.int32)
        self.lib.orc_get_cover_state(self.h, _ptr(cov, c_u8p), _ptr(con, c_i32p), _ptr(cid, c_i32p), ctypes.c_int64(n))
        return cov, con, cid

    def grid(self, with_bits=True):
        origin = np.zeros(3, np.float64); dims = np.zeros(3, np.int32)
        self.lib.orc_get_grid(self.h, _ptr(origin, c_f64p), _ptr(dims, c_i32p), None, ctypes.c_int64(0))
        bits = None
        if with_bits:
            g = int(dims[0]) * int(dims[1]) * int(dims[2])
            bits = np.zeros((g + 7) // 8, np.uint8)
            self.lib.orc_get_grid(self.h, None, None, _ptr(bits, c_u8p), ctypes.c_int64(len(bits)))
        return origin, dims, bits

    def trace_enable(self, on=True):
        self.lib.orc_trace_enable(self.h, ctypes.c_int32(1 if on else 0))

    def trace(self):
        """[(stage, vp_id, pose5, row_words)] in execution order."""
        n = self.lib.orc_trace_count(self.h)
        words = self.lib.orc_trace_words(self.h)
        out = []
        for i in range(n):
            st = ctypes.c_int32(); vid = ctypes.c_int32()
            pose = np.zeros(5, np.float64); row = np.zeros(words, np.uint32)
            self.lib.orc_trace_get(self.h, ctypes.c_int64(i), ctypes.byref(st), ctypes.byref(vid), _ptr(pose, c_f64p), _ptr(row, c_u32p))
            out.append((st.value, vid.value, pose, row))
        return out

    def prune_order(self):
        n = self.lib.orc_prune_order(self.h, None, ctypes.c_int64(0))
        out = np.zeros(n, np.int32)
        self.lib.orc_prune_order(self.h, _ptr(out, c_i32p), ctypes.c_int64(n))
        return out

    def counters(self):
        out = np.zeros(8, np.int64)
        self.lib.orc_get_counters(self.h, _ptr(out, c_i64p))
        return out

    def stage_seconds(self):
        out = np.zeros(5, np.float64)
        self.lib.orc_get_stage_seconds(self.h, _ptr(out, c_f64p))
        return out

    def evaluate(self, stage, pose5, restrict_rows=None, threads=1):
        """Evaluator pass over a batch of poses -> (count[n], rows[n, words], counters[8])."""
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        n = len(pose5)
        words = (self.P + 31) // 32
        count = np.zeros(n, np.int32); rows = np.zeros((n, words), np.uint32); ctr = np.zeros(8, np.int64)
        if restrict_rows is not None:
            restrict_rows = np.ascontiguousarray(restrict_rows, dtype=np.uint32)
        rc = self.lib.orc_evaluate(self.h, ctypes.c_int32(stage), _ptr(pose5, c_f64p), ctypes.c_int64(n), _ptr(restrict_rows, c_u32p),
                                   _ptr(count, c_i32p), _ptr(rows, c_u32p), ctypes.c_int32(threads), _ptr(ctr, c_i64p))
        if rc != 0:
            raise RuntimeError(f"orc_evaluate failed: {rc}")
        return count, rows, ctr

    def safety_check(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        out = np.zeros(len(xyz), np.uint8)
        self.lib.orc_safety_check(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)), _ptr(out, c_u8p))
        return out.astype(bool)


# -- micro known-answer taps -----------------------------------------------------------------
def ray_ids(res, origin, start, end, cap=4096, alt=False):
    lib = load(alt)
    o = np.asarray(origin, np.float64); s = np.asarray(start, np.float64); e = np.asarray(end, np.float64)
    out = np.zeros((cap, 3), np.int32)
    k = lib.orc_ray_ids(ctypes.c_double(res), _ptr(o, c_f64p), _ptr(s, c_f64p), _ptr(e, c_f64p), _ptr(out, c_i32p), ctypes.c_int64(cap))
    return out[:min(k, cap)].copy()


def biray_ids(res, origin, start, end, cap=4096, alt=False):
    lib = load(alt)
    o = np.asarray(origin, np.float64); s = np.asarray(start, np.float64); e = np.asarray(end, np.float64)
    out = np.zeros((cap, 6), np.int32)
    k = lib.orc_biray_ids(ctypes.c_double(res), _ptr(o, c_f64p), _ptr(s, c_f64p), _ptr(e, c_f64p), _ptr(out, c_i32p), ctypes.c_int64(cap))
    return out[:min(k, cap)].copy()


def fov_normals(params, pitch, yaw, alt=False):
    lib = load(alt)
    out = np.zeros(12, np.float64)
    lib.orc_fov_normals(ctypes.byref(params), ctypes.c_double(pitch), ctypes.c_double(yaw), _ptr(out, c_f64p))
    return out.reshape(4, 3)


def inside_fov(params, pose5, pts, alt=False):
    lib = load(alt)
    pose5 = np.ascontiguousarray(pose5, np.float64); pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    out = np.zeros(len(pts), np.uint8)
    lib.orc_inside_fov(ctypes.byref(params), _ptr(pose5, c_f64p), _ptr(pts, c_f64p), ctypes.c_int64(len(pts)), _ptr(out, c_u8p))
    return out.astype(bool)


def intbound(s, ds):
    return load().orc_intbound(s, ds)


def pitch_yaw(v):
    v = np.ascontiguousarray(v, np.float64); out = np.zeros(2, np.float64)
    load().orc_pitch_yaw(_ptr(v, c_f64p), _ptr(out, c_f64p))
    return out


# -- oracle/_ref: the reference's own raycast.cpp / perception_utils.cpp (built from /root/reference
#    by the Makefile against shim/ headers; see ref_capi.cpp) --------------------------------------
_REF = None


def ref_path():
    return os.path.join(_HERE, "_ref", "libfcvm_ref.so")


def ref_available():
    return os.path.exists(ref_path())


def ref_load():
    global _REF
    if _REF is None:
        lib = ctypes.CDLL(ref_path())
        for fn in ("ref_mod", "ref_intbound"):
            getattr(lib, fn).restype = ctypes.c_double
        lib.ref_signum.restype = ctypes.c_int
        lib.ref_signum.argtypes = [ctypes.c_int]
        lib.ref_mod.argtypes = [ctypes.c_double, ctypes.c_double]
        lib.ref_intbound.argtypes = [ctypes.c_double, ctypes.c_double]
        lib.ref_ray_ids.restype = ctypes.c_int64
        lib.ref_biray_ids.restype = ctypes.c_int64
        lib.ref_create.restype = ctypes.c_void_p
        lib.ref_create.argtypes = [ctypes.c_void_p]
        lib.ref_destroy.argtypes = [ctypes.c_void_p]
        _REF = lib
    return _REF


def ref_ray_ids(res, origin, start, end, cap=4096, bi=False):
    lib = ref_load()
    o = np.asarray(origin, np.float64); s = np.asarray(start, np.float64); e = np.asarray(end, np.float64)
    w = 6 if bi else 3
    out = np.zeros((cap, w), np.int32)
    fn = lib.ref_biray_ids if bi else lib.ref_ray_ids
    k = fn(ctypes.c_double(res), _ptr(o, c_f64p), _ptr(s, c_f64p), _ptr(e, c_f64p), _ptr(out, c_i32p), ctypes.c_int64(cap), ctypes.c_int64(cap))
    return out[:min(k, cap)].copy()


def ref_fov_normals(params, pitch, yaw):
    out = np.zeros(12, np.float64)
    ref_load().ref_fov_normals(ctypes.byref(params), ctypes.c_double(pitch), ctypes.c_double(yaw), _ptr(out, c_f64p))
    return out.reshape(4, 3)


def ref_inside_fov(params, pose5, pts):
    pose5 = np.ascontiguousarray(pose5, np.float64); pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    out = np.zeros(len(pts), np.uint8)
    ref_load().ref_inside_fov(ctypes.byref(params), _ptr(pose5, c_f64p), _ptr(pts, c_f64p), ctypes.c_int64(len(pts)), _ptr(out, c_u8p))
    return out.astype(bool)


class RefEvaluator:
    """Evaluator passes E1..E4 run by the reference's RayCaster + PerceptionUtils objects."""

    def __init__(self, params):
        self.lib = ref_load()
        self.params = params
        self.h = ctypes.c_void_p(self.lib.ref_create(ctypes.byref(params)))
        self.P = 0

    def __del__(self):
        try:
            if self.h:
                self.lib.ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def setMapPointCloud(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.lib.ref_set_map_cloud(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def setModel(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.P = len(xyz)
        self.lib.ref_set_model(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def evaluate(self, stage, pose5, restrict_rows=None, threads=1, want_rows=True):
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        n = len(pose5)
        words = (self.P + 31) // 32
        count = np.zeros(n, np.int32); rows = np.zeros((n, words), np.uint32) if want_rows else None; ctr = np.zeros(4, np.int64)
        if restrict_rows is not None:
            restrict_rows = np.ascontiguousarray(restrict_rows, dtype=np.uint32)
        rc = self.lib.ref_evaluate_mt(self.h, ctypes.c_int32(stage), _ptr(pose5, c_f64p), ctypes.c_int64(n), _ptr(restrict_rows, c_u32p),
                                      _ptr(count, c_i32p), _ptr(rows, c_u32p), _ptr(ctr, c_i64p), ctypes.c_int32(threads))
        if rc != 0:
            raise RuntimeError(f"ref_evaluate failed: {rc}")
        return count, rows, ctr


# -- oracle/_ref: the reference's own viewpoint_manager.cpp (ref_vm_capi.cpp) -----------------------
_REFVM = None


def refvm_path():
    return os.path.join(_HERE, "_ref", "libfcvm_refvm.so")


def refvm_available():
    return os.path.exists(refvm_path())


def refvm_load():
    global _REFVM
    if _REFVM is None:
        lib = ctypes.CDLL(refvm_path())
        lib.refvm_create.restype = ctypes.c_void_p
        lib.refvm_create.argtypes = [ctypes.c_void_p]
        lib.refvm_destroy.argtypes = [ctypes.c_void_p]
        for fn in ("refvm_get_viewpoints", "refvm_get_uncovered", "refvm_get_vps_pose", "refvm_get_vps_voxcount", "refvm_get_vps_contri",
                   "refvm_get_cover_state", "refvm_lookups"):
            getattr(lib, fn).restype = ctypes.c_int64
        _REFVM = lib
    return _REFVM


class RefViewpointManager:
    """predrecon::ViewpointManager compiled from the reference's own viewpoint_manager.cpp, driven
    through the same call sequence as `Oracle`.  ONE instance per process at a time: the injected
    seed counters (clock shim, RandomSample shim) are process-wide."""

    def __init__(self, params):
        self.lib = refvm_load()
        self.params = params
        self.h = ctypes.c_void_p(self.lib.refvm_create(ctypes.byref(params)))
        self.P = 0

    def __del__(self):
        try:
            if self.h:
                self.lib.refvm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self):
        self.lib.refvm_reset(self.h)

    def setMapPointCloud(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        return self.lib.refvm_set_map_cloud(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def setModel(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.P = len(xyz)
        return self.lib.refvm_set_model(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def setNormals(self, pts, normals):
        pts = np.ascontiguousarray(pts, dtype=np.float64); normals = np.ascontiguousarray(normals, dtype=np.float64)
        return self.lib.refvm_set_normals(self.h, _ptr(pts, c_f64p), _ptr(normals, c_f64p), ctypes.c_int64(len(pts)))

    def setInitViewpoints(self, pos_dir):
        pos_dir = np.ascontiguousarray(pos_dir, dtype=np.float32)
        return self.lib.refvm_set_init_viewpoints(self.h, _ptr(pos_dir, c_f32p), ctypes.c_int64(len(pos_dir)))

    def updateViewpoints(self):
        return self.lib.refvm_update(self.h)

    def getUpdatedViewpoints(self):
        n = self.lib.refvm_get_viewpoints(self.h, None, None, None, ctypes.c_int64(0))
        ids = np.zeros(n, np.int32); pose = np.zeros((n, 5), np.float64); vox = np.zeros(n, np.int32)
        self.lib.refvm_get_viewpoints(self.h, _ptr(ids, c_i32p), _ptr(pose, c_f64p), _ptr(vox, c_i32p), ctypes.c_int64(n))
        return ids, pose, vox

    def getUncoveredModelCloud(self):
        n = self.lib.refvm_get_uncovered(self.h, None, ctypes.c_int64(0))
        out = np.zeros((n, 3), np.float32)
        self.lib.refvm_get_uncovered(self.h, _ptr(out, c_f32p), ctypes.c_int64(n))
        return out

    def vps_pose(self):
        n = self.lib.refvm_get_vps_pose(self.h, None, ctypes.c_int64(0))
        out = np.zeros((n, 5), np.float64)
        self.lib.refvm_get_vps_pose(self.h, _ptr(out, c_f64p), ctypes.c_int64(n))
        return out

    def vps_voxcount(self):
        n = self.lib.refvm_get_vps_voxcount(self.h, None, ctypes.c_int64(0))
        out = np.zeros(n, np.int32)
        self.lib.refvm_get_vps_voxcount(self.h, _ptr(out, c_i32p), ctypes.c_int64(n))
        return out

    def vps_contri(self):
        n = self.lib.refvm_get_vps_contri(self.h, None, ctypes.c_int64(0))
        out = np.zeros(n, np.int32)
        self.lib.refvm_get_vps_contri(self.h, _ptr(out, c_i32p), ctypes.c_int64(n))
        return out

    def cover_state(self):
        n = self.lib.refvm_get_cover_state(self.h, None, None, None, ctypes.c_int64(0))
        cov = np.zeros(n, np.uint8); con = np.zeros(n, np.int32); cid = np.zeros(n, np.int32)
        self.lib.refvm_get_cover_state(self.h, _ptr(cov, c_u8p), _ptr(con, c_i32p), _ptr(cid, c_i32p), ctypes.c_int64(n))
        return cov, con, cid

    def lookups(self):
        s = ctypes.c_int64(0)
        occ = self.lib.refvm_lookups(self.h, ctypes.byref(s))
        return int(occ), int(s.value)


# -- pin-hole coverage evaluation (fcvm_oracle_cov.hpp; hcplanner.cpp:1032-1211) -------------------
class CameraParams(ctypes.Structure):
    """hcplanner/{fx, fy, cx, cy} + viewpoint_manager/{fov_h, fov_w}; same layout as fcvm_camera in include/fcvm.h."""
    _fields_ = [("fx", ctypes.c_double), ("fy", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double),
                ("fov_h", ctypes.c_double), ("fov_w", ctypes.c_double)]


class CoverageOracle:
    def __init__(self, params, camera, alt=False):
        self.lib = load(alt)
        self.params, self.camera = params, camera
        self.h = ctypes.c_void_p(self.lib.orc_cov_create(ctypes.byref(params), ctypes.byref(camera)))
        self.n = 0

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_cov_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def setCloud(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        self.n = len(xyz)
        self.lib.orc_cov_set_cloud(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    def image(self):
        wh = np.zeros(2, np.int32); res = np.zeros(2, np.float64)
        self.lib.orc_cov_image(self.h, _ptr(wh, c_i32p), _ptr(res, c_f64p))
        return int(wh[0]), int(wh[1]), float(res[0]), float(res[1])

    def project(self, pose5, pts):
        pose5 = np.ascontiguousarray(pose5, np.float64); pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        dist = np.zeros(len(pts), np.float64); xy = np.zeros((len(pts), 2), np.int32)
        self.lib.orc_cov_project(self.h, _ptr(pose5, c_f64p), _ptr(pts, c_f64p), ctypes.c_int64(len(pts)), _ptr(dist, c_f64p), _ptr(xy, c_i32p))
        return dist, xy

    def pinhole(self, pose5, threads=1):
        """PinHoleCamera per pose -> (offsets[m+1], ids, covered[n] union, counters[3])."""
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        m = len(pose5)
        offsets = np.zeros(m + 1, np.int64); covered = np.zeros(self.n, np.uint8); ctr = np.zeros(3, np.int64)
        total = self.lib.orc_cov_pinhole(self.h, _ptr(pose5, c_f64p), ctypes.c_int64(m), _ptr(offsets, c_i64p), None, ctypes.c_int64(0), None,
                                         ctypes.c_int32(threads), None)
        ids = np.zeros(total, np.int32)
        self.lib.orc_cov_pinhole(self.h, _ptr(pose5, c_f64p), ctypes.c_int64(m), _ptr(offsets, c_i64p), _ptr(ids, c_i32p), ctypes.c_int64(total),
                                 _ptr(covered, c_u8p), ctypes.c_int32(threads), _ptr(ctr, c_i64p))
        return offsets, ids, covered, ctr

    def CoverageEvaluation(self, pose5):
        """-> (coverage rate, group_offsets[m+1], group_ids, ever_visible[n])."""
        pose5 = np.ascontiguousarray(pose5, dtype=np.float64).reshape(-1, 5)
        m = len(pose5)
        rate = ctypes.c_double(0); offsets = np.zeros(m + 1, np.int64); ever = np.zeros(self.n, np.uint8)
        cap = max(1, self.n * 4)
        while True:
            ids = np.zeros(cap, np.int32)
            total = self.lib.orc_cov_evaluate(self.h, _ptr(pose5, c_f64p), ctypes.c_int64(m), ctypes.byref(rate), _ptr(offsets, c_i64p),
                                              _ptr(ids, c_i32p), ctypes.c_int64(cap), _ptr(ever, c_u8p))
            if total <= cap:
                return rate.value, offsets, ids[:total], ever
            cap = int(total)


# -- planner-grid loops (fcvm_oracle_plan.hpp; hcplanner.cpp:658-692, 805-863; hcsolver.cpp:556-583) ----------
c_i8p = ctypes.POINTER(ctypes.c_int8)


class PlanOracle:
    """origin[3], resolution, voxel_num[3] + the three voxel buffers of the planner's SDFMap (x-major int8 arrays or None)."""

    def __init__(self, origin, resolution, voxel_num, hc=None, inflate=None, internal=None):
        self.lib = load()
        self.origin = np.ascontiguousarray(origin, np.float64)
        self.dims = np.ascontiguousarray(voxel_num, np.int32)
        bufs = [None if b is None else np.ascontiguousarray(b, np.int8).ravel() for b in (hc, inflate, internal)]
        self._bufs = bufs
        self.h = ctypes.c_void_p(self.lib.orc_plan_create(_ptr(self.origin, c_f64p), ctypes.c_double(resolution), _ptr(self.dims, c_i32p),
                                                          *[_ptr(b, c_i8p) for b in bufs]))

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def lightStateDet(self, cand, rosa):
        cand = np.ascontiguousarray(cand, np.float64).reshape(-1, 3); rosa = np.ascontiguousarray(rosa, np.float32).reshape(-1, 3)
        n = len(cand)
        det = np.zeros(n, np.uint8); cross = np.zeros(n, np.int32); near = np.zeros(n, np.int32)
        self.lib.orc_plan_light_state(self.h, _ptr(cand, c_f64p), ctypes.c_int64(n), _ptr(rosa, c_f32p), ctypes.c_int64(len(rosa)), _ptr(det, c_u8p),
                                      _ptr(cross, c_i32p), _ptr(near, c_i32p))
        return det, cross, near

    def safetyBox(self, vps, safe_radius):
        vps = np.ascontiguousarray(vps, np.float32).reshape(-1, 3)
        ok = np.zeros(len(vps), np.uint8)
        self.lib.orc_plan_safety_box(self.h, _ptr(vps, c_f32p), ctypes.c_int64(len(vps)), ctypes.c_double(safe_radius), _ptr(ok, c_u8p))
        return ok

    def lineOfSight(self, pts, threads=1):
        pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
        n = len(pts)
        safe = np.zeros((n, n), np.uint8); dist = np.zeros((n, n), np.float64)
        self.lib.orc_plan_line_of_sight(self.h, _ptr(pts, c_f64p), ctypes.c_int64(n), _ptr(safe, c_u8p), _ptr(dist, c_f64p), ctypes.c_int32(threads))
        return safe, dist

    def counters(self):
        out = np.zeros(2, np.int64)
        self.lib.orc_plan_counters(self.h, _ptr(out, c_i64p))
        return out

    # ---- SDFMap::SetOcc / InternalSpace / OuterCheck (sdf_map.cpp:263-473) on the three buffers
    def setOcc(self, xyz):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self.lib.orc_plan_set_occ(self.h, _ptr(xyz, c_f32p), ctypes.c_int64(len(xyz)))

    SetOcc = setOcc

    def InternalSpace(self, seg_clouds, seg_ends, inflate_num, proj_interval, thickness):
        """seg_clouds: list of float32 [n_s, 3]; seg_ends: [nseg, 6] = (start xyz, end xyz) of each skeleton segment."""
        xyz = np.ascontiguousarray(np.concatenate(seg_clouds) if len(seg_clouds) else np.zeros((0, 3)), np.float32)
        offs = np.concatenate([[0], np.cumsum([len(c) for c in seg_clouds])]).astype(np.int64)
        ends = np.ascontiguousarray(seg_ends, np.float64).reshape(-1, 6)
        self.lib.orc_plan_internal_space(self.h, _ptr(xyz, c_f32p), _ptr(offs, c_i64p), ctypes.c_int64(len(seg_clouds)), _ptr(ends, c_f64p),
                                         ctypes.c_int32(inflate_num), ctypes.c_double(proj_interval), ctypes.c_double(thickness))

    def OuterCheck(self, outers, check_scale):
        o = np.ascontiguousarray(outers, np.float64).reshape(-1, 6)
        self.lib.orc_plan_outer_check(self.h, _ptr(o, c_f64p), ctypes.c_int64(len(o)), ctypes.c_double(check_scale))

    def buffers(self):
        G = int(np.prod(self.dims.astype(np.int64)))
        hc = np.zeros(G, np.int8); inf = np.zeros(G, np.int8); internal = np.zeros(G, np.int8)
        self.lib.orc_plan_get_buffers(self.h, _ptr(hc, c_i8p), _ptr(inf, c_i8p), _ptr(internal, c_i8p))
        return hc, inf, internal

    def oob_writes(self):
        self.lib.orc_plan_oob_writes.restype = ctypes.c_int64
        return int(self.lib.orc_plan_oob_writes(self.h))


class RefPlanGrid:
    """The planner's call-site loops driving the REFERENCE's RayCaster (oracle/_ref/libfcvm_ref.so)."""

    def __init__(self, origin, resolution, voxel_num, hc=None, inflate=None, internal=None):
        self.lib = ref_load()
        self.lib.ref_plan_grid_create.restype = ctypes.c_void_p
        self.lib.ref_plan_grid_destroy.argtypes = [ctypes.c_void_p]
        o = np.ascontiguousarray(origin, np.float64); d = np.ascontiguousarray(voxel_num, np.int32)
        self._bufs = [None if b is None else np.ascontiguousarray(b, np.int8).ravel() for b in (hc, inflate, internal)]
        self.h = ctypes.c_void_p(self.lib.ref_plan_grid_create(_ptr(o, c_f64p), ctypes.c_double(resolution), _ptr(d, c_i32p),
                                                               *[_ptr(b, c_i8p) for b in self._bufs]))

    def __del__(self):
        try:
            if self.h:
                self.lib.ref_plan_grid_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def light_state_cross(self, vp, vertex, max_calls=100000):
        vp = np.ascontiguousarray(vp, np.float64); vertex = np.ascontiguousarray(vertex, np.float64)
        cap = ctypes.c_int32(0)
        c = self.lib.ref_light_state_cross(self.h, _ptr(vp, c_f64p), _ptr(vertex, c_f64p), ctypes.c_int64(max_calls), ctypes.byref(cap))
        return int(c), bool(cap.value)

    def cast_marks(self, a, b, cap=4096, max_calls=100000):
        """voxels the InternalSpace / OuterCheck marking loop touches for the ray a -> b -> ([k, 3] map indices, capped?)"""
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
        out = np.zeros((cap, 3), np.int32)
        c = ctypes.c_int32(0)
        self.lib.ref_cast_marks.restype = ctypes.c_int64
        k = self.lib.ref_cast_marks(self.h, _ptr(a, c_f64p), _ptr(b, c_f64p), _ptr(out, c_i32p), ctypes.c_int64(cap), ctypes.c_int64(max_calls), ctypes.byref(c))
        return out[:min(int(k), cap)].copy(), bool(c.value)

    def straight_line(self, p1, p2, max_calls=100000):
        p1 = np.ascontiguousarray(p1, np.float64); p2 = np.ascontiguousarray(p2, np.float64)
        cap = ctypes.c_int32(0)
        s = self.lib.ref_straight_line(self.h, _ptr(p1, c_f64p), _ptr(p2, c_f64p), ctypes.c_int64(max_calls), ctypes.byref(cap))
        return bool(s), bool(cap.value)
