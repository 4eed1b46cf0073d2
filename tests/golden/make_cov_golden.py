#!/usr/bin/env python
"""Generates tests/golden/cov_<scene>.npz: the pin of the pin-hole coverage evaluation.

Runs HERE only (reads /root/reference).  The reference ships, for MBS / pipe / Christ, the OUTPUT of its own
`CoverageEvaluation` run (hierarchical_coverage_planner/solution/Traj/):
    TrajInfo<Scene>.txt    one trajectory pose per line (x, y, z, pitch, yaw; printed with 6 decimals) = `TrajPose`
    CloudInfo<Scene>.txt   per pose, the points of `visible_group[k]` (hcopp.cpp:108-109, hctraj.cpp:197-221)
evaluated against `Fullmodel` = the launch file's `hcplanner/fullcloud` (.pcd; hcplanner.cpp:83-84, :252).

The fixture holds the poses, the reference's groups mapped back to cloud indices (by the 6-decimal text of the
coordinates), and what the CPU oracle (oracle/fcvm_oracle_cov.hpp) produces for the same poses.  The cloud itself is
`map_cloud` of tests/golden/<scene>.npz.  The oracle reproduces the shipped groups up to the poses whose 6-decimal
rounding moves a point across a pixel / FOV boundary (MBS 82 581 of 82 589 ids; counts recorded in the fixture).

    python tests/golden/make_cov_golden.py [scene ...]
"""
import os
import re
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fc_planner_b200 import launch_camera, launch_params, scenes  # noqa: E402
from oracle import oracle  # noqa: E402

REF = "/root/reference/FC-Planner/src/hierarchical_coverage_planner"
SCENES = {"mbs": ("MBS", "MBSmore.pcd"), "pipe": ("Pipe", "pipe.pcd"), "christ": ("Christ", "christmore.pcd")}
NUM = re.compile(r"-?\d+\.\d+")


def text_key(p):
    return tuple("%.6f" % float(v) for v in p)      # std::to_string(double) prints %f


def main():
    for scene in (sys.argv[1:] or list(SCENES)):
        name, pcd = SCENES[scene]
        cloud = scenes.read_pcd_xyz(os.path.join(REF, "data", pcd))
        z = np.load(os.path.join(HERE, f"{scene}.npz"))
        assert np.array_equal(z["map_cloud"], cloud), "the scene fixture holds another cloud"
        traj = np.array([[float(x) for x in NUM.findall(l)] for l in open(os.path.join(REF, "solution", "Traj", f"TrajInfo{name}.txt"))])
        poses = np.ascontiguousarray(traj[:, 1:6])
        lines = open(os.path.join(REF, "solution", "Traj", f"CloudInfo{name}.txt")).read().split("\n")
        index = {}
        for i, p in enumerate(cloud):
            index.setdefault(text_key(p), []).append(i)
        ref_offsets, ref_ids, unknown = [0], [], 0
        for k in range(len(poses)):
            g = np.array([float(x) for x in NUM.findall(lines[2 * k + 1])]).reshape(-1, 3)
            ids = set()
            for p in g:
                hit = index.get(text_key(p))
                if hit is None:
                    unknown += 1
                else:
                    ids.update(hit)
            ref_ids.extend(sorted(ids))
            ref_offsets.append(len(ref_ids))
        ref_offsets, ref_ids = np.array(ref_offsets, np.int64), np.array(ref_ids, np.int32)

        prm, cam = launch_params(scene), launch_camera(scene)
        ce = oracle.CoverageOracle(prm, cam)
        ce.setCloud(cloud)
        rate, offs, ids, ever = ce.CoverageEvaluation(poses)
        p_offs, p_ids, covered, ctr = ce.pinhole(poses, threads=8)
        same = only_ref = only_orc = 0
        for k in range(len(poses)):
            a, b = set(ids[offs[k]:offs[k + 1]].tolist()), set(ref_ids[ref_offsets[k]:ref_offsets[k + 1]].tolist())
            same += len(a & b); only_ref += len(b - a); only_orc += len(a - b)
        print(f"{scene}: {len(poses)} poses x {len(cloud)} points; reference groups {len(ref_ids)} ids ({unknown} unmatched points); "
              f"oracle {len(ids)}; same {same}, only reference {only_ref}, only oracle {only_orc}; rate {rate:.6f}")
        np.savez_compressed(os.path.join(HERE, f"cov_{scene}.npz"), scene=scene, poses=poses, ref_offsets=ref_offsets, ref_ids=ref_ids,
                            agreement=np.array([same, only_ref, only_orc, unknown], np.int64), rate=np.float64(rate), group_offsets=offs,
                            group_ids=ids, ever=np.packbits(ever, bitorder="little"), pin_offsets=p_offs, pin_ids_crc=np.uint32(zlib.crc32(p_ids.tobytes())),
                            covered=np.packbits(covered, bitorder="little"), counters=ctr)


if __name__ == "__main__":
    main()
