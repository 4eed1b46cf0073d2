#!/usr/bin/env python
"""How far do the rays of the bench scene's E1 pass really march?  Runs the pass with the two measurement builds
(FCVM_DEFS=EVAL_STATS=1 / 2 / 3, see k_eval) and prints, per ray: biNextId calls of the reference, lane-steps marched by
the kernel, rays ended by the in-flight free-box certificate / by an occupied voxel.
    FCVM_LIB=<variant.so> python profiles/march_stats.py <1|2|3> [views]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fc_planner_b200 import launch_params, scenes  # noqa: E402
from fc_planner_b200.manager import ViewpointManager  # noqa: E402

mode = int(sys.argv[1])
views = int(sys.argv[2]) if len(sys.argv) > 2 else 25000
pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
prm = launch_params("synthetic", seed=0, resolution=0.1)
vps = scenes.plant_viewpoints(pts, nrm, prims, views, prm.viewpoints_distance, seed=100)
poses = scenes.poses_from_viewpoints(vps)
vm = ViewpointManager(prm)
vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm)
c0 = np.array(vm.counters())
cnt, _ = vm.evaluate(1, poses, want_rows=False)
c = np.array(vm.counters()) - c0
out = {"mode": mode, "views": views, "rays": int(c[1]), "lookups": int(c[2]), "visible": int(np.sum(cnt)),
       "calls_per_ray": float(c[2]) / 2 / max(1, int(c[1]))}
if mode == 1:
    out.update(lane_steps=int(c[3]), ended_by_certificate=int(c[4]), lane_steps_per_ray=float(c[3]) / max(1, int(c[1])))
elif mode == 2:
    out.update(calls_saved_by_certificate=int(c[3]), ended_by_occupied_voxel=int(c[4]))
else:
    out.update(blocked_at_first_call_after_loading=int(c[3]), blocked_within_3_calls_of_loading=int(c[4]))
print(json.dumps(out))
