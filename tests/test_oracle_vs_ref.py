"""Pins the CPU oracle against the reference's OWN source: oracle/_ref is raycast.cpp and
perception_utils.cpp compiled from /root/reference (oracle/Makefile, oracle/ref_capi.cpp) against
stand-in Eigen / ROS headers.  Bar: bit-exact (integer voxel sequences, boolean verdicts, and FP64
values produced by the same scalar operations).

The library is built here (where /root/reference exists) and travels to the GPU box as a built
file; the tests skip when it is absent.
"""
import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


def test_intbound_mod_signum_match_reference_source():
    lib = oracle.ref_load(); orc = oracle.load()
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.uniform(-500, 500, 4000), rng.integers(-50, 50, 500).astype(float),
                        [0.0, -0.0, 1e-300, -1e-300, 0.5, -0.5, 1 - 2**-53, -(1 - 2**-53), 123456.999999999, -7.000000000000001]])
    ds = np.concatenate([rng.integers(-40, 41, len(s) - 6).astype(float), [0.0, -0.0, 1.0, -1.0, 3.0, -3.0]])
    for a, b in zip(s, ds):
        r, o = lib.ref_intbound(a, b), orc.orc_intbound(a, b)
        assert r == o or (np.isnan(r) and np.isnan(o)), (a, b, r, o)
        assert lib.ref_mod(a, 1.0) == orc.orc_mod(a, 1.0)
    for x in (-3, 0, 2):
        assert lib.ref_signum(x) == (0 if x == 0 else (-1 if x < 0 else 1))


def _rays(n, seed, res):
    """Random, axis-aligned, exactly diagonal, lattice-snapped and degenerate rays."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-20, 20, (n, 3)); b = a + rng.normal(0, 6, (n, 3))
    k = n // 6
    b[:k, 1:] = a[:k, 1:]                                     # axis aligned (two idle axes: NaN tDelta)
    b[k:2 * k, 2] = a[k:2 * k, 2]                             # planar
    a[2 * k:3 * k] = np.round(a[2 * k:3 * k] / res) * res     # start on lattice corners
    b[2 * k:3 * k] = a[2 * k:3 * k] + np.round(rng.normal(0, 8, (k, 3))) * res   # exact diagonals / ties
    b[3 * k:3 * k + 5] = a[3 * k:3 * k + 5]                   # start == end
    a[3 * k + 5:4 * k] = a[3 * k + 5:4 * k].astype(np.float32)   # float-rounded like PCL points
    return a, b


def _compare_marchers(res, origin, a, b, bi):
    """Same voxel sequence, call for call.  The reference marcher has no step cap and really does
    run away on some rays that start exactly on a lattice plane (a finished axis ties with a live
    one at t = 1 and wins): there the harness stops the reference after CAP calls, the oracle must
    have stopped at its Manhattan + 8 cap, and the two sequences must agree up to that point.
    Returns the number of such rays."""
    CAP, runaway = 4096, 0
    for s, e in zip(a, b):
        r = oracle.ref_ray_ids(res, origin, s, e, cap=CAP, bi=bi)
        o = oracle.biray_ids(res, origin, s, e, cap=CAP) if bi else oracle.ray_ids(res, origin, s, e, cap=CAP)
        if len(r) == CAP:
            runaway += 1
            assert len(o) < CAP and np.array_equal(r[:len(o)], o), (s, e)
        else:
            assert r.shape == o.shape and np.array_equal(r, o), (s, e)
    return runaway


@pytest.mark.parametrize("res", [0.4, 0.1])
def test_one_directional_marcher_visits_same_voxels(res):
    origin = np.array([-31.7, -29.3, -33.1])
    a, b = _rays(1500, 1, res)
    runaway = _compare_marchers(res, origin, a, b, bi=False)
    assert runaway < 0.02 * len(a)


@pytest.mark.parametrize("res", [0.4, 0.1])
def test_bidirectional_marcher_visits_same_voxels(res):
    origin = np.array([-31.7, -29.3, -33.1])
    a, b = _rays(1500, 2, res)
    runaway = _compare_marchers(res, origin, a, b, bi=True)
    assert runaway < 0.02 * len(a)


def test_fov_normals_and_inside_fov_match_reference_source(small_plant):
    prm = small_plant["params"]
    rng = np.random.default_rng(3)
    for _ in range(200):
        pitch, yaw = rng.uniform(-1.5, 1.3), rng.uniform(-3.2, 3.2)
        assert np.array_equal(oracle.ref_fov_normals(prm, pitch, yaw), oracle.fov_normals(prm, pitch, yaw))
    for _ in range(40):
        pose = np.concatenate([rng.uniform(-5, 5, 3), [rng.uniform(-1.4, 1.2), rng.uniform(-3.1, 3.1)]])
        pts = pose[:3] + rng.normal(0, 6, (4000, 3))
        # points on the range sphere and (nearly) on the FOV planes
        n = oracle.fov_normals(prm, pose[3], pose[4])
        d = rng.normal(0, 1, (500, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        on_sphere = pose[:3] + d * prm.max_dist
        on_plane = pose[:3] + np.cross(n[rng.integers(0, 4, 500)], d) * rng.uniform(0.5, 9, (500, 1))
        allp = np.concatenate([pts, on_sphere, on_plane])
        assert np.array_equal(oracle.ref_inside_fov(prm, pose, allp), oracle.inside_fov(prm, pose, allp))


@pytest.mark.parametrize("stage", [1, 2, 3, 4])
def test_evaluator_masks_match_reference_objects(small_plant, stage):
    """E1..E4 visibility masks: oracle vs the reference's RayCaster + PerceptionUtils driven by the
    call-site loops of viewpoint_manager.cpp; also the algorithmic counters used by the roofline."""
    from fc_planner_b200 import scenes
    prm = small_plant["params"]
    orc = oracle.Oracle(prm); ref = oracle.RefEvaluator(prm)
    for m in (orc, ref):
        m.setMapPointCloud(small_plant["points"]); m.setModel(small_plant["points"])
    poses = scenes.poses_from_viewpoints(small_plant["viewpoints"][:120])
    rng = np.random.default_rng(5)
    poses[:, :3] += rng.normal(0, 0.3, (len(poses), 3))
    poses[:, 3:] += rng.normal(0, 0.2, (len(poses), 2))
    restrict = None
    if stage == 2:
        _, restrict, _ = orc.evaluate(1, poses)
        poses = poses.copy(); poses[:, 3:] += 0.05
    oc, orows, octr = orc.evaluate(stage, poses, restrict_rows=restrict)
    rc, rrows, rctr = ref.evaluate(stage, poses, restrict_rows=restrict)
    assert np.array_equal(orows, rrows) and np.array_equal(oc, rc)
    assert oc.sum() > 300
    assert np.array_equal(octr[:4], rctr)     # pair tests, rays, voxel lookups, walk lookups
