#!/usr/bin/env python
"""Coverage evaluation at the scale of BASELINE.json configs[4]: the 1 000 000-point synthetic plant as the full cloud,
m trajectory-like poses (candidate viewpoints of the plant); GPU through the public call vs the single-threaded oracle on
a prefix of the poses (the sweep across poses is sequential, so a prefix is itself a valid run).
    python profiles/cov_scale.py [poses] [oracle_prefix]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fc_planner_b200 import launch_camera, launch_params, scenes  # noqa: E402
from fc_planner_b200.coverage import CoverageEvaluator  # noqa: E402
from oracle import oracle  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 150
pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
prm, cam = launch_params("synthetic", resolution=0.1), launch_camera("mbs")
poses = scenes.poses_from_viewpoints(scenes.plant_viewpoints(pts, nrm, prims, m, prm.viewpoints_distance, seed=7))
ce = CoverageEvaluator(prm, cam)
ce.setCloud(pts)
ce.CoverageEvaluation(poses[:64])
c0 = ce.counters()
t = time.perf_counter(); rate, offs, ids, ever = ce.CoverageEvaluation(poses); wall = time.perf_counter() - t
c = ce.counters() - c0
scan = ce.last_scan_ms()
orc = oracle.CoverageOracle(prm, cam)
orc.setCloud(pts)
t = time.perf_counter(); o_rate, o_offs, o_ids, o_ever = orc.CoverageEvaluation(poses[:k]); cpu = time.perf_counter() - t
g = ce.CoverageEvaluation(poses[:k])
b_alg = 16.0 * c[0] + 20.0 * c[1] + 2.0 * m * (len(pts) / 8.0)
print(json.dumps({"points": len(pts), "poses": m, "gpu_wall_ms": 1e3 * wall, "kernels_ms": ce.last_kernel_ms() if False else None, "scan_kernel_ms": scan,
                  "coverage_rate": rate, "group_ids": int(len(ids)), "pair_tests": int(c[0]), "projections": int(c[1]),
                  "chunks_scanned_frac": float(c[4]) / (m * ((len(pts) + 1023) // 1024)), "shared_pixel_points": int(c[5]),
                  "pair_tests_per_s": c[0] / (scan * 1e-3), "algorithmic_gb_per_s": b_alg / (scan * 1e-3) / 1e9,
                  "oracle_prefix_poses": k, "oracle_prefix_s": cpu, "oracle_full_estimate_s": cpu * m / k,
                  "prefix_identical": bool(g[0] == o_rate and np.array_equal(g[1], o_offs) and np.array_equal(g[2], o_ids) and np.array_equal(g[3], o_ever))}))
