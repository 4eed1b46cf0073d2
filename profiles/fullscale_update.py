#!/usr/bin/env python
"""One full call sequence at the scale of BASELINE.json configs[4] per GPU share: P = 1 000 000 points,
25 000 initial candidate viewpoints, 0.1 m grid, on one B200 (information only: the CPU oracle would
need hours here, so there is no CPU number beside it; parity at this size is covered by
tests/test_gpu_fullsize.py).  FCVM_PROFILE=1 prints the split of the wall time.
    FCVM_PROFILE=1 python profiles/fullscale_update.py [views]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fc_planner_b200 import launch_params, scenes  # noqa: E402
from fc_planner_b200.manager import ViewpointManager  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
prm = launch_params("synthetic", seed=0, resolution=0.1)
vps = scenes.plant_viewpoints(pts, nrm, prims, views, prm.viewpoints_distance, seed=100)
vm = ViewpointManager(prm)
t0 = time.perf_counter(); vm.setMapPointCloud(pts); t_map = time.perf_counter() - t0
t0 = time.perf_counter(); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm); t_model = time.perf_counter() - t0
t0 = time.perf_counter(); vm.setInitViewpoints(vps); t_init = time.perf_counter() - t0
t0 = time.perf_counter(); rc = vm.updateViewpoints(); t_upd = time.perf_counter() - t0
assert rc == 0
c = vm.counters()
print(json.dumps({"P": len(pts), "V_initial": len(vps), "setMapPointCloud_s": t_map, "setModel_setNormals_s": t_model, "setInitViewpoints_s": t_init,
                  "updateViewpoints_s": t_upd, "final_viewpoints": len(vm.getUpdatedViewpoints()), "uncovered_points": int(len(vm.getUncoveredModelCloud())),
                  "candidates_total": int(len(vm.vps_pose())), "pair_tests": int(c[0]), "rays": int(c[1]), "voxel_lookups": int(c[2]), "cap_trips": int(c[4]),
                  "kernel_launches": int(c[5])}))
