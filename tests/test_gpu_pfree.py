"""k_eval's half-ray certificate (summed-area table over 8^3-voxel cells: a viewpoint-side half-ray whose voxel box is empty
is counted, not marched) must change nothing observable: masks, per-viewpoint counts and the reference's look-up / ray /
pair-test counts are those of the fully marching kernel (FCVM_NO_PFREE=1) and of the CPU oracle, for all four evaluators."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _manager(prm, **env):
    from fc_planner_b200.manager import ViewpointManager
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return ViewpointManager(prm)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


@pytest.mark.parametrize("res,scale,npts", [(0.2, 0.5, 40000), (0.1, 0.35, 60000)])
def test_certified_half_rays_change_nothing(oracle_mod, res, scale, npts):
    from fc_planner_b200 import launch_params, scenes
    pts, nrm, prims = scenes.make_plant(n_points=npts, scale=scale, seed=4)
    prm = launch_params("synthetic", seed=4, resolution=res)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 600, prm.viewpoints_distance, seed=4)
    poses = scenes.poses_from_viewpoints(vps)
    # a few awkward poses, checked separately below: inside the structure's bounding box, far outside the map, and ON voxel lattice
    # planes - there the reference's marcher runs past its midpoint voxel and such rays are never certified
    extra = poses[:6].copy()
    extra[0, :3] = pts[0].astype(np.float64) + 0.05
    extra[1, :3] += 40.0
    extra[2, :3] = np.round(extra[2, :3] / res) * res
    extra[3, 0] = np.round(extra[3, 0] / res) * res
    a, b = _manager(prm), _manager(prm, FCVM_NO_PFREE=1)
    orc = oracle_mod.Oracle(prm)
    for m in (a, b, orc):
        m.setMapPointCloud(pts); m.setModel(pts)
    for stage in (1, 3, 4):
        ca0, cb0 = a.counters(), b.counters()
        cnt_a, rows_a = a.evaluate(stage, poses)
        cnt_b, rows_b = b.evaluate(stage, poses)
        ocnt, orows, oc = orc.evaluate(stage, poses)
        assert np.array_equal(rows_a, rows_b) and np.array_equal(cnt_a, cnt_b)
        assert np.array_equal(rows_a[:, :orows.shape[1]], orows) and np.array_equal(cnt_a, ocnt)
        da, db = a.counters() - ca0, b.counters() - cb0
        assert np.array_equal(da[:5], db[:5]), (stage, da, db)                       # pair tests, rays, look-ups, walk look-ups, cap trips
        assert da[1] == oc[1] and da[2] == oc[2] and da[3] == oc[3] and da[4] == 0   # ... and they are the reference's counts
        if stage == 1:
            e1 = rows_a
        # the awkward poses: masks as the oracle's (a march that overruns its midpoint is "not visible" on both sides); look-up and
        # trip counts are not compared with the oracle there (the reference would march on to the map's edge)
        ca0, cb0 = a.counters(), b.counters()
        xa, ra = a.evaluate(stage, extra)
        xb, rb = b.evaluate(stage, extra)
        xo, ro, xc = orc.evaluate(stage, extra)
        assert np.array_equal(ra, rb) and np.array_equal(ra[:, :ro.shape[1]], ro) and np.array_equal(xa, xo)
        assert (a.counters() - ca0)[4] == (b.counters() - cb0)[4] and ((a.counters() - ca0)[4] > 0) == (xc[4] > 0)
    cnt_a, rows_a = a.evaluate(2, poses, restrict_rows=e1)
    cnt_b, rows_b = b.evaluate(2, poses, restrict_rows=e1)
    ocnt, orows, oc = orc.evaluate(2, poses, restrict_rows=e1[:, :orows.shape[1]])
    assert np.array_equal(rows_a, rows_b) and np.array_equal(rows_a[:, :orows.shape[1]], orows) and np.array_equal(cnt_a, ocnt)
    assert cnt_a.sum() > 1000
