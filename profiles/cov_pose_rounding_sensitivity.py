#!/usr/bin/env python
"""Is the residue between the coverage oracle and the reference's shipped CloudInfo*.txt (MBS: 8 ids only in the reference, 7
only in ours, of 82 589) explained by the 6-decimal printing of the trajectory poses?  The reference evaluated the poses it
had in memory and printed them with `%f`; the fixture can only re-read the printed text.  If the explanation holds, moving
every pose component by a random amount within the printing error (+-5e-7) must change `visible_group` by about as many ids
as the residue, and the ids of the residue must be among those that flip under such moves.  CPU only (the oracle).
    python profiles/cov_pose_rounding_sensitivity.py [scene] [trials]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN, load_golden  # noqa: E402
from fc_planner_b200 import launch_camera, launch_params  # noqa: E402
from oracle import oracle  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "mbs"
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 6
_, zs = load_golden(scene)
z = np.load(os.path.join(GOLDEN, f"cov_{scene}.npz"))
poses = z["poses"]


def groups(p):
    oc = oracle.CoverageOracle(launch_params(scene), launch_camera(scene))
    oc.setCloud(zs["map_cloud"])
    _, offs, ids, _ = oc.CoverageEvaluation(p)
    return {(k, int(i)) for k in range(len(p)) for i in ids[offs[k]:offs[k + 1]]}


base = groups(poses)
ref = {(k, int(i)) for k in range(len(poses)) for i in z["ref_ids"][z["ref_offsets"][k]:z["ref_offsets"][k + 1]]}
residue = base ^ ref
rng = np.random.default_rng(0)
flips, sizes = set(), []
for t in range(trials):
    moved = groups(poses + rng.uniform(-5e-7, 5e-7, poses.shape))
    d = base ^ moved
    sizes.append(len(d))
    flips |= d
print(json.dumps({"scene": scene, "poses": int(len(poses)), "ids_oracle": len(base), "ids_reference": len(ref),
                  "residue_vs_reference": len(residue), "only_reference": len(ref - base), "only_ours": len(base - ref),
                  "trials": trials, "ids_changed_by_a_move_within_the_printing_error": sizes,
                  "residue_ids_that_flip_in_some_trial": len(residue & flips),
                  "residue_points_that_flip_in_some_trial": len({i for _, i in residue} & {i for _, i in flips})}))
