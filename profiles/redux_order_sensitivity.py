#!/usr/bin/env python
"""The stated-epsilon count of north_star / SURVEY.md A.2.  Eigen is absent from this image, so its 3-term reduction order
under SSE2 is an assumption: (a0*b0 + a1*b1) + a2*b2 (default oracle) - the alternative a0*b0 + (a1*b1 + a2*b2) is built as
oracle/_build/libfcvm_oracle_alt.so.  This script runs both oracles on the shipped scenes (tests/golden/*.npz inputs) and on
a prefix of the bench batch and counts what differs: pair tests whose E1 verdict changes (they can only be points within an
ulp of a FOV plane or of max_dist), and whether the selected viewpoint set changes.  CPU only.
    python profiles/redux_order_sensitivity.py [--bench-views 256] > profiles/r2/redux_order_sensitivity.json"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fc_planner_b200 import launch_params, scenes  # noqa: E402
from oracle import oracle  # noqa: E402


def run_sequence(orc, z):
    orc.reset(); orc.setMapPointCloud(z["map_cloud"]); orc.setModel(z["model"]); orc.setNormals(z["model"].astype(np.float64), z["normals"])
    assert orc.updateViewpoints() == 0
    return orc.getUpdatedViewpoints(), orc.vps_pose()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bench-views", type=int, default=256)
    args = ap.parse_args()
    oracle.build()
    out = {"assumption": "Eigen SSE2 3-term reduction = (a0*b0 + a1*b1) + a2*b2; alternative = a0*b0 + (a1*b1 + a2*b2)", "scenes": {}}
    for scene in ("mbs", "pipe", "christ", "pipe4x"):
        z = np.load(os.path.join(ROOT, "tests", "golden", f"{scene}.npz"))
        prm = launch_params({"pipe4x": "pipe"}.get(scene, scene), seed=int(z["seed"]))
        a, b = oracle.Oracle(prm), oracle.Oracle(prm, alt=True)
        t = time.perf_counter()
        (ia, pa, xa), va = run_sequence(a, z)
        (ib, pb, xb), vb = run_sequence(b, z)
        # E1 of every initial candidate pose: verdict-level difference
        poses = z["vps_pose"]
        for o in (a, b):
            o.setMapPointCloud(z["map_cloud"]); o.setModel(z["model"])
        ca, ra, ka = a.evaluate(1, poses); cb, rb, kb = b.evaluate(1, poses)
        diff_bits = int(np.unpackbits((ra ^ rb).view(np.uint8)).sum())
        out["scenes"][scene] = {
            "P": int(len(z["model"])), "candidate_poses": int(len(poses)), "pair_tests": int(ka[0]),
            "redux_order_sensitive_pairs": diff_bits,
            "fov_pass_count_differs": bool(ka[1] != kb[1]), "rays_default": int(ka[1]), "rays_alternative": int(kb[1]),
            "selected_set_identical": bool(np.array_equal(ia, ib) and np.array_equal(xa, xb)),
            "selected_poses_bit_identical": bool(np.array_equal(pa, pb)),
            "candidate_poses_bit_identical": bool(va.shape == vb.shape and np.array_equal(va, vb)),
            "max_abs_pose_difference": float(np.abs(va - vb).max()) if va.shape == vb.shape else None,
            "seconds": time.perf_counter() - t}
    pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
    prm = launch_params("synthetic", seed=0, resolution=0.1)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 25000, prm.viewpoints_distance, seed=100)
    poses = scenes.poses_from_viewpoints(vps)[:args.bench_views]
    a, b = oracle.Oracle(prm), oracle.Oracle(prm, alt=True)
    for o in (a, b):
        o.setMapPointCloud(pts); o.setModel(pts)
    threads = len(os.sched_getaffinity(0))
    ca, ra, ka = a.evaluate(1, poses, threads=threads); cb, rb, kb = b.evaluate(1, poses, threads=threads)
    out["bench_prefix"] = {"workload": f"bench scene (P = 10^6, 0.1 m), first {len(poses)} viewpoints, E1", "pair_tests": int(ka[0]),
                           "redux_order_sensitive_pairs": int(np.unpackbits((ra ^ rb).view(np.uint8)).sum()),
                           "rays_default": int(ka[1]), "rays_alternative": int(kb[1])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
