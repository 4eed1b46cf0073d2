"""Host-side mirror of the planner's pin-hole coverage evaluation over the C ABI (include/fcvm.h,
`fcvm_cov_*`): `hierarchical_coverage_planner::PinHoleCamera` (hcplanner.cpp:1128-1194) and
`::CoverageEvaluation` (:1032-1126), batched over all poses of a path / trajectory.

Same names and argument meaning as the reference; clouds and pose sets are numpy arrays.  The CUDA
library is the only implementation: a missing device raises, nothing degrades to the CPU.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, f32p, f64p, i32p, i64p, ptr, u8p


class CoverageEvaluator:
    def __init__(self, params, camera):
        """params: perception_utils/* (FOV half-angles, max_dist) + device; camera: hcplanner/{fx,fy,cx,cy} + fov."""
        self.lib = _lib.load()
        self.params, self.camera = params, camera
        self.h = ctypes.c_void_p(self.lib.fcvm_cov_create(ctypes.byref(params), ctypes.byref(camera)))
        if not self.h:
            raise _lib.FcvmError(-1, self.lib.fcvm_last_error().decode())
        self.n = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.fcvm_cov_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setCloud(self, cloud):
        """`Fullmodel` (hcplanner.cpp:252) / `checkCloud` (:1129)."""
        xyz = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
        check(self.lib.fcvm_cov_set_cloud(self.h, ptr(xyz, f32p), len(xyz)))
        self.n = len(xyz)

    def _poses(self, poseSet):
        return np.ascontiguousarray(poseSet, dtype=np.float64).reshape(-1, 5)

    def PinHoleCamera(self, poseSet):
        """-> (offsets[m+1], covered ids ascending per pose, CoverState[n] of the path loop :253-261)."""
        pose5 = self._poses(poseSet)
        m = len(pose5)
        offsets = np.zeros(m + 1, np.int64)
        covered = np.zeros(self.n, np.uint8)
        total = check(self.lib.fcvm_cov_pinhole(self.h, ptr(pose5, f64p), m, ptr(offsets, i64p), None, 0, ptr(covered, u8p)))
        return offsets, self._fetch(total), covered

    def _fetch(self, total):
        ids = np.zeros(int(total), np.int32)
        assert check(self.lib.fcvm_cov_fetch_ids(self.h, ptr(ids, i32p), len(ids))) == total
        return ids

    def path_coverage(self, poseSet):
        """The path-coverage loop of hcplanner.cpp:250-271 -> (coverage, CoverState[n])."""
        pose5 = self._poses(poseSet)
        covered = np.zeros(self.n, np.uint8)
        check(self.lib.fcvm_cov_pinhole(self.h, ptr(pose5, f64p), len(pose5), None, None, 0, ptr(covered, u8p)))
        return float(covered.sum()) / float(self.n), covered

    def CoverageEvaluation(self, poseSet):
        """-> (coverageRate, group_offsets[m+1], group ids, ever_visible[n]); group k = the points pushed into
        `visible_group[k]` (:1102-1108)."""
        pose5 = self._poses(poseSet)
        m = len(pose5)
        rate = ctypes.c_double(0.0)
        offsets = np.zeros(m + 1, np.int64)
        ever = np.zeros(self.n, np.uint8)
        total = check(self.lib.fcvm_cov_evaluate(self.h, ptr(pose5, f64p), m, ctypes.byref(rate), ptr(offsets, i64p), None, 0, ptr(ever, u8p)))
        return rate.value, offsets, self._fetch(total), ever

    def chunk_points(self):
        """points per culling chunk (counters()[4] counts chunks scanned)"""
        return int(self.lib.fcvm_cov_chunk_points())

    def counters(self):
        """[pair tests, projections, out-of-image, kernels launched, chunks scanned, duplicate-pixel slow paths]"""
        out = np.zeros(6, np.int64)
        check(self.lib.fcvm_cov_counters(self.h, ptr(out, i64p)))
        return out

    def last_kernel_ms(self):
        return float(self.lib.fcvm_cov_last_kernel_ms(self.h))

    def last_scan_ms(self):
        return float(self.lib.fcvm_cov_last_scan_ms(self.h))
