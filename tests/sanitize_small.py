"""Small end-to-end run for memory-safety checking: all four evaluator stages, the full call
sequence, lattice-aligned viewpoints (stray rays) and viewpoints outside the map (literal marcher).
    FCVM_LIB=fc-planner_b200/_build/libfcvm_dbg.so python tests/sanitize_small.py     (index-checking build)
    compute-sanitizer --tool memcheck python tests/sanitize_small.py                  (where the tool is open)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fc_planner_b200 import launch_params, scenes  # noqa: E402
from fc_planner_b200.manager import ViewpointManager  # noqa: E402

pts, nrm, prims = scenes.make_plant(n_points=3000, scale=0.3, seed=1)
prm = launch_params("mbs", z_ground=0, seed=3)
vps = scenes.plant_viewpoints(pts, nrm, prims, 96, prm.viewpoints_distance, seed=1)
vm = ViewpointManager(prm)
vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm)
poses = scenes.poses_from_viewpoints(vps)
tot = 0
for stage in (1, 3, 4):
    cnt, rows = vm.evaluate(stage, poses)
    tot += int(cnt.sum())
_, rows1 = vm.evaluate(1, poses)
cnt2, _ = vm.evaluate(2, poses, restrict_rows=rows1)
lat = poses.copy(); lat[:, :3] = np.round(lat[:, :3] / prm.resolution) * prm.resolution
vm.evaluate(1, lat)
far = poses.copy(); far[:, :3] += 40.0          # outside the padded map: literal marcher + out-of-map lookups
vm.evaluate(4, far)
vm.evaluate_csr(3, poses)
vm.setInitViewpoints(vps)
assert vm.updateViewpoints() == 0
print("bounds violations", vm.debug_bounds_violations())
print("sanitize_small ok:", tot, int(cnt2.sum()), len(vm.getUpdatedViewpoints()), "final viewpoints, counters", vm.counters().tolist())
