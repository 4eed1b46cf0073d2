"""Host-side mirror of three loops of the planner that run on its own SDFMap, over the C ABI (`fcvm_grid_*`):
`hierarchical_coverage_planner::lightStateDet` (hcplanner.cpp:805-863), the candidate safety box (:658-676) and the
straight-line test of `HCSolver::search_Path` (hcsolver.cpp:556-583), each for whole batches.  CUDA only."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, f32p, f64p, i8p, i32p, i64p, ptr, u8p


class PlannerGrid:
    def __init__(self, origin, resolution, voxel_num, occupancy_hc=None, occupancy_inflate=None, occupancy_internal=None, device=0):
        """map_origin_, resolution_, map_voxel_num_ + the three voxel buffers (x-major int8 arrays of prod(voxel_num), or None)."""
        self.lib = _lib.load()
        o = np.ascontiguousarray(origin, np.float64)
        d = np.ascontiguousarray(voxel_num, np.int32)
        bufs = [None if b is None else np.ascontiguousarray(b, np.int8).ravel() for b in (occupancy_hc, occupancy_inflate, occupancy_internal)]
        for b in bufs:
            if b is not None and b.size != int(d[0]) * int(d[1]) * int(d[2]):
                raise ValueError("voxel buffer size does not match voxel_num")
        self.h = ctypes.c_void_p(self.lib.fcvm_grid_create(device, ptr(o, f64p), float(resolution), ptr(d, i32p), *[ptr(b, i8p) for b in bufs]))
        if not self.h:
            raise _lib.FcvmError(-1, self.lib.fcvm_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.fcvm_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def lightStateDet(self, candidates, rosa_cloud):
        """-> (det[n] 1 = outside, crossCount[n], nearest vertex id[n])"""
        c = np.ascontiguousarray(candidates, np.float64).reshape(-1, 3)
        r = np.ascontiguousarray(rosa_cloud, np.float32).reshape(-1, 3)
        det = np.zeros(len(c), np.uint8); cross = np.zeros(len(c), np.int32); near = np.zeros(len(c), np.int32)
        check(self.lib.fcvm_grid_light_state(self.h, ptr(c, f64p), len(c), ptr(r, f32p), len(r), ptr(det, u8p), ptr(cross, i32p), ptr(near, i32p)))
        return det, cross, near

    def safetyBox(self, viewpoints, safe_radius):
        v = np.ascontiguousarray(viewpoints, np.float32).reshape(-1, 3)
        ok = np.zeros(len(v), np.uint8)
        check(self.lib.fcvm_grid_safety_box(self.h, ptr(v, f32p), len(v), float(safe_radius), ptr(ok, u8p)))
        return ok

    def lineOfSight(self, positions, want_dist=True):
        """search_Path's straight-line test for all ordered pairs -> (safe[n, n], dist[n, n])"""
        p = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        n = len(p)
        safe = np.zeros((n, n), np.uint8)
        dist = np.zeros((n, n), np.float64) if want_dist else None
        check(self.lib.fcvm_grid_line_of_sight(self.h, ptr(p, f64p), n, ptr(safe, u8p), ptr(dist, f64p)))
        return safe, dist

    def counters(self):
        out = np.zeros(3, np.int64)
        check(self.lib.fcvm_grid_counters(self.h, ptr(out, i64p)))
        return out

    def last_kernel_ms(self):
        return float(self.lib.fcvm_grid_last_kernel_ms(self.h))
