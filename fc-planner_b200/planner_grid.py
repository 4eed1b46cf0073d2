"""Host-side mirror of three loops of the planner that run on its own SDFMap, over the C ABI (`fcvm_grid_*`):
`hierarchical_coverage_planner::lightStateDet` (hcplanner.cpp:805-863), the candidate safety box (:658-676) and the
straight-line test of `HCSolver::search_Path` (hcsolver.cpp:556-583), each for whole batches - and of the SDFMap calls that
build that map's buffers (`SetOcc`, `InternalSpace`, `OuterCheck`, plan_env/src/sdf_map.cpp:263-473), same names.  CUDA only."""
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
        self.dims = d
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

    # ---- building the buffers on the device (start from PlannerGrid(origin, res, voxel_num): all three zero) ----
    def SetOcc(self, cloud):
        """SDFMap::SetOcc (sdf_map.cpp:440-455) / initHCMap's marking loop (:178-187)."""
        c = np.ascontiguousarray(cloud, np.float32).reshape(-1, 3)
        check(self.lib.fcvm_grid_set_occ(self.h, ptr(c, f32p), len(c)))

    def InternalSpace(self, seg_clouds, seg_ends, inflate_num, proj_interval, thickness):
        """SDFMap::InternalSpace (sdf_map.cpp:263-395).  seg_clouds: the per-segment clouds (list of float32 [n_s, 3]);
        seg_ends[s] = (start xyz, end xyz) of the segment's two skeleton vertices."""
        xyz = np.ascontiguousarray(np.concatenate(seg_clouds), np.float32)
        offs = np.concatenate([[0], np.cumsum([len(c) for c in seg_clouds])]).astype(np.int64)
        ends = np.ascontiguousarray(seg_ends, np.float64).reshape(-1, 6)
        assert len(ends) == len(seg_clouds)
        check(self.lib.fcvm_grid_internal_space(self.h, ptr(xyz, f32p), ptr(offs, i64p), len(seg_clouds), ptr(ends, f64p), int(inflate_num),
                                                float(proj_interval), float(thickness)))

    def OuterCheck(self, outers, check_scale):
        """SDFMap::OuterCheck (sdf_map.cpp:457-473): outers = n x (position, direction)."""
        o = np.ascontiguousarray(outers, np.float64).reshape(-1, 6)
        check(self.lib.fcvm_grid_outer_check(self.h, ptr(o, f64p), len(o), float(check_scale)))

    def buffers(self):
        """(occupancy_buffer_hc_, occupancy_inflate_buffer_hc_, occupancy_buffer_internal_) as x-major int8 arrays."""
        G = int(np.prod(self.dims.astype(np.int64)))
        out = [np.zeros(G, np.int8) for _ in range(3)]
        check(self.lib.fcvm_grid_get_buffers(self.h, *[ptr(b, i8p) for b in out]))
        return tuple(out)

    def oob_writes(self):
        return int(self.lib.fcvm_grid_oob_writes(self.h))

    def counters(self):
        out = np.zeros(3, np.int64)
        check(self.lib.fcvm_grid_counters(self.h, ptr(out, i64p)))
        return out

    def last_kernel_ms(self):
        return float(self.lib.fcvm_grid_last_kernel_ms(self.h))
