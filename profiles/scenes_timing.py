#!/usr/bin/env python
"""BASELINE.json configs[0..3]: the shipped scenes (MBS, pipe, Christ the Redeemer, and the pipe-4x substitute for the
absent "Pacifico" scene) through the reference's call sequence reset -> setMapPointCloud -> setModel -> setNormals ->
updateViewpoints (viewpoints generated from the normals), GPU library vs the single-threaded CPU oracle in the same run;
the selected viewpoint set, poses and counts must be identical.  Inputs: tests/golden/<scene>.npz.
    python profiles/scenes_timing.py [scene ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from fc_planner_b200.manager import ViewpointManager  # noqa: E402
from oracle import oracle  # noqa: E402


def sequence(m, z):
    m.reset()
    t0 = time.perf_counter(); m.setMapPointCloud(z["map_cloud"]); t_map = time.perf_counter() - t0
    m.setModel(z["model"]); m.setNormals(z["model"].astype(np.float64), z["normals"])
    t0 = time.perf_counter(); rc = m.updateViewpoints(); t_upd = time.perf_counter() - t0
    assert rc == 0
    return t_map, t_upd


for scene in (sys.argv[1:] or ["mbs", "pipe", "christ", "pipe4x"]):
    prm, z = load_golden(scene)
    vm = ViewpointManager(prm)
    first = sequence(vm, z)
    gi, gp, gx = vm.final_arrays()
    c = vm.counters()
    warm = min((sequence(vm, z) for _ in range(3)), key=lambda t: t[1])
    orc = oracle.Oracle(prm)
    cpu = sequence(orc, z)
    oi, op, ox = orc.getUpdatedViewpoints()
    print(json.dumps({"scene": scene, "map_points": int(len(z["map_cloud"])), "model_points": int(len(z["model"])),
                      "candidates": int(len(vm.vps_pose())), "final_viewpoints": int(len(gi)), "rays": int(c[1]), "voxel_lookups": int(c[2]),
                      "kernel_launches": int(c[5]), "gpu_update_ms_first_call": 1e3 * first[1], "gpu_update_ms": 1e3 * warm[1],
                      "gpu_set_map_ms": 1e3 * warm[0], "cpu_update_ms": 1e3 * cpu[1], "cpu_set_map_ms": 1e3 * cpu[0], "cpu_cores": 1, "cpu_kind": "port",
                      "speedup_update": cpu[1] / warm[1],
                      "identical_first_run": bool(np.array_equal(gi, oi) and np.array_equal(gp, op) and np.array_equal(gx, ox)),
                      "matches_golden": bool(np.array_equal(gi, z["final_ids"]) and np.array_equal(gp, z["final_pose"]))}))
