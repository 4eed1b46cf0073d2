import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def small_plant():
    """6 000-point synthetic plant at 0.4 scale, 0.4 m grid, MBS parameters without ground."""
    from fc_planner_b200 import launch_params, scenes
    pts, nrm, prims = scenes.make_plant(n_points=6000, scale=0.4, seed=0)
    prm = launch_params("mbs", z_ground=0, seed=7)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 400, prm.viewpoints_distance, seed=0)
    return dict(points=pts, normals=nrm, prims=prims, params=prm, viewpoints=vps)


def rows_to_sets(rows):
    return [np.flatnonzero(np.unpackbits(r.view(np.uint8), bitorder="little")) for r in rows]


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(scene):
    """Committed fixture of one shipped scene: inputs + everything the oracle produced for the
    8-call sequence (tests/golden/make_golden.py)."""
    from fc_planner_b200 import launch_params
    z = np.load(os.path.join(GOLDEN, f"{scene}.npz"))
    prm = launch_params({"pipe4x": "pipe"}.get(scene, scene), seed=int(z["seed"]))   # pipe4x: configs[3] substitute, see make_golden.py
    return prm, z


def run_sequence(mgr, z, trace=True):
    """reset -> setMapPointCloud -> setModel -> setNormals -> updateViewpoints (viewpoints generated
    from the normals, viewpoint_manager.cpp:145-171) on a manager-like object."""
    if trace and len(z["trace_stage"]):
        mgr.trace_enable()
    mgr.reset()
    mgr.setMapPointCloud(z["map_cloud"])
    mgr.setModel(z["model"])
    mgr.setNormals(z["model"].astype(np.float64), z["normals"])
    assert mgr.updateViewpoints() == 0


def compare_with_golden(mgr, z, final, check_trace=True):
    ids, pose, vox = final
    assert np.array_equal(ids, z["final_ids"])                 # selected viewpoint set, reference order
    assert np.array_equal(pose, z["final_pose"])               # FP64 bit-exact
    assert np.array_equal(vox, z["final_vox"])
    assert np.array_equal(mgr.vps_pose(), z["vps_pose"])
    gv, ov = mgr.vps_voxcount(), z["vps_voxcount"]
    n = min(len(gv), len(ov))
    assert np.array_equal(gv[:n], ov[:n]) and not gv[n:].any() and not ov[n:].any()
    assert np.array_equal(mgr.vps_contri(), z["vps_contri"])
    cov, con, cid = mgr.cover_state()
    assert np.array_equal(cov, z["covered"]) and np.array_equal(con, z["cover_contrib"]) and np.array_equal(cid, z["contrib_id"])
    assert np.array_equal(mgr.getUncoveredModelCloud(), z["uncovered"])
    c = mgr.counters()
    assert c[4] == 0                                            # no marcher cap trips
    assert c[1] == z["counters"][1] and c[2] == z["counters"][2] and c[3] == z["counters"][3]   # rays, lookups, walk lookups
    if check_trace and len(z["trace_stage"]):
        tr = mgr.trace()
        W = z["trace_rows"].shape[1]
        assert len(tr) == len(z["trace_stage"])
        key = lambda st, vp: (int(st), int(vp))
        want = {}
        for i in range(len(z["trace_stage"])):
            want.setdefault(key(z["trace_stage"][i], z["trace_vp"][i]), []).append(i)
        got = {}
        for st, vp, p, row in tr:
            got.setdefault(key(st, vp), []).append((p, row))
        assert got.keys() == want.keys()
        bad = 0
        for k, idxs in want.items():
            assert len(got[k]) == len(idxs), k
            for (p, row), i in zip(got[k], idxs):
                bad += int(not np.array_equal(row[:W], z["trace_rows"][i])) + int(not np.array_equal(p, z["trace_pose"][i]))
        assert bad == 0
