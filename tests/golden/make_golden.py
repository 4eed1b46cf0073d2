#!/usr/bin/env python
"""Generates the golden fixtures tests/golden/<scene>.npz from the reference's shipped scenes.

Runs HERE only (it reads /root/reference, which does not exist on the GPU box); the fixtures and
this script are committed.  For each scene (MBS, pipe, christ):

  inputs   map cloud (the launch file's `hcplanner/fullcloud`), model cloud = voxel-centroid
           down-sample of the model source at `hcplanner/model_downsample_size`, ground-filtered
           like hcplanner.cpp:128-140 and snapped to the nearest source point and restricted to points whose own voxel is occupied in the
           map grid (hcplanner.cpp:150-158); one outward normal per model point.
           ROSA + cutting-plane normals (the reference's own source of these inputs) are out of
           scope and wall-clock seeded, so normals come from a documented deterministic rule: PCA
           of the 24 nearest map points, oriented away from the centroid of the map points within
           10 m (a stand-in for the skeleton axis the reference measures from, hcplanner.cpp:956-973).
  outputs  everything the CPU oracle produces for the 8-call sequence with those inputs and the
           launch file's parameters: final viewpoints, all candidate poses, counts, cover state,
           uncovered cloud, prune order, counters, and the E1..E4 visibility masks of every
           evaluator call.

    python tests/golden/make_golden.py [scene ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fc_planner_b200 import launch_params, scenes  # noqa: E402
from fc_planner_b200.params import model_leaf  # noqa: E402
from oracle import oracle  # noqa: E402

REF = "/root/reference/FC-Planner/src"
SCENES = {
    "mbs": dict(map="hierarchical_coverage_planner/data/MBSmore.pcd", model="hierarchical_coverage_planner/data/MBS.pcd", ground=-10.5),
    "pipe": dict(map="hierarchical_coverage_planner/data/pipe.pcd", model="hierarchical_coverage_planner/data/pipe.pcd", ground=None),
    "christ": dict(map="hierarchical_coverage_planner/data/christmore.pcd", model="rosa/data/christ.pcd", ground=None),
    # BASELINE.json configs[3] names a "Pacifico" scene the reference does not ship (SURVEY.md preface): substitute =
    # pipe with the model leaf halved (2.0 -> 1.0 m), one candidate per model point => ~4x the candidate viewpoints
    "pipe4x": dict(map="hierarchical_coverage_planner/data/pipe.pcd", model="hierarchical_coverage_planner/data/pipe.pcd", ground=None,
                   params="pipe", leaf=1.0, trace=False),
}


def build_inputs(scene):
    from scipy.spatial import cKDTree
    cfg = SCENES[scene]
    prm = launch_params(cfg.get("params", scene), seed=20231019)
    map_cloud = scenes.read_pcd_xyz(os.path.join(REF, cfg["map"]))
    src = scenes.read_pcd_xyz(os.path.join(REF, cfg["model"]))
    if cfg["ground"] is not None:
        src = src[src[:, 2] > cfg["ground"]]
    model = scenes.voxel_centroids(src, cfg.get("leaf") or model_leaf(scene))
    # snap every centroid to the nearest point of the model source, so that thin surfaces keep their
    # points (the reference keeps a centroid only if its own voxel is occupied, hcplanner.cpp:150-158)
    model = src[cKDTree(src.astype(np.float64)).query(model.astype(np.float64), k=1)[1]]
    # keep points whose own voxel is occupied in the HC grid of the map cloud
    orc = oracle.Oracle(prm)
    orc.setMapPointCloud(map_cloud)
    origin, dims, bits = orc.grid()
    idx = np.floor((model.astype(np.float64) - origin) * (1 / prm.resolution)).astype(np.int64)
    inb = np.all((idx >= 0) & (idx < dims), axis=1)
    adr = (idx[:, 0] * dims[1] + idx[:, 1]) * dims[2] + idx[:, 2]
    occ = np.zeros(len(model), bool)
    occ[inb] = (bits[adr[inb] >> 3] >> (adr[inb] & 7)) & 1
    model = np.ascontiguousarray(model[occ])
    _, first = np.unique(model, axis=0, return_index=True)
    model = np.ascontiguousarray(model[np.sort(first)])
    # normals
    tree = cKDTree(map_cloud.astype(np.float64))
    _, nn = tree.query(model.astype(np.float64), k=24)
    normals = np.zeros((len(model), 3))
    for i in range(len(model)):
        q = map_cloud[nn[i]].astype(np.float64)
        c = q - q.mean(0)
        w, v = np.linalg.eigh(c.T @ c)
        n = v[:, 0]
        big = map_cloud[tree.query_ball_point(model[i].astype(np.float64), 10.0)].astype(np.float64)
        out = model[i].astype(np.float64) - big.mean(0)
        if np.dot(n, out) < 0:
            n = -n
        normals[i] = n / np.linalg.norm(n)
    return prm, map_cloud, model, normals


def run_oracle(prm, map_cloud, model, normals, init_vps=None, trace=True):
    orc = oracle.Oracle(prm)
    if trace:
        orc.trace_enable()
    orc.reset()
    orc.setMapPointCloud(map_cloud)
    orc.setModel(model)
    orc.setNormals(model.astype(np.float64), normals)
    if init_vps is not None:
        orc.setInitViewpoints(init_vps)
    rc = orc.updateViewpoints()
    assert rc == 0
    ids, pose, vox = orc.getUpdatedViewpoints()
    tr = orc.trace()
    cov, con, cid = orc.cover_state()
    return dict(
        final_ids=ids, final_pose=pose, final_vox=vox, vps_pose=orc.vps_pose(), vps_voxcount=orc.vps_voxcount(),
        vps_contri=orc.vps_contri(), covered=cov, cover_contrib=con, contrib_id=cid, uncovered=orc.getUncoveredModelCloud(),
        prune_order=orc.prune_order(), counters=orc.counters(),
        trace_stage=np.array([t[0] for t in tr], np.int8), trace_vp=np.array([t[1] for t in tr], np.int32),
        trace_pose=np.array([t[2] for t in tr], np.float64).reshape(-1, 5),
        trace_rows=np.array([t[3] for t in tr], np.uint32).reshape(len(tr), -1) if tr else np.zeros((0, 0), np.uint32))


def main():
    names = sys.argv[1:] or list(SCENES)
    for scene in names:
        prm, map_cloud, model, normals = build_inputs(scene)
        out = run_oracle(prm, map_cloud, model, normals, trace=SCENES[scene].get("trace", True))
        path = os.path.join(HERE, f"{scene}.npz")
        np.savez_compressed(path, scene=scene, seed=np.uint64(prm.seed), map_cloud=map_cloud, model=model, normals=normals, **out)
        c = out["counters"]
        print(f"{scene}: map {len(map_cloud)} pts, model P={len(model)}, candidates V={len(out['vps_pose'])}, final {len(out['final_ids'])}, "
              f"uncovered {len(out['uncovered'])}, evaluator calls {len(out['trace_stage'])}, rays {c[1]}, lookups {c[2]}, cap trips {c[4]} "
              f"-> {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
