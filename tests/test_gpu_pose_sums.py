"""The device pose-sum pipeline (fcvm_pose.cu: correctly rounded asin / acos + flagged rays to the host's libm, ordered sums
on the device) and the arg-max ownership (k_own_best / k_own_apply) against the round-1 forms of the same steps - every
ray through glibc on the host (FCVM_HOST_SUMS=1), ordered row sweep (FCVM_OWN_SWEEP=1) - and against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _manager(prm, **env):
    from fc_planner_b200.manager import ViewpointManager
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return ViewpointManager(prm)          # the A/B switches are read when the handle is created
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _run(vm, pts, nrm, vps):
    vm.reset(); vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm)
    vm.setInitViewpoints(vps)
    pose_init, vox_init, contri_init = vm.vps_pose(), vm.vps_voxcount(), vm.vps_contri()
    cov_init = vm.cover_state()
    assert vm.updateViewpoints() == 0
    return dict(pose_init=pose_init, vox_init=vox_init, contri_init=contri_init, cov_init=cov_init, final=vm.final_arrays(), pose=vm.vps_pose(),
                vox=vm.vps_voxcount(), contri=vm.vps_contri(), cov=vm.cover_state(), unc=vm.getUncoveredModelCloud())


def _same(a, b):
    for k in a:
        x, y = a[k], b[k]
        if isinstance(x, tuple):
            assert all(np.array_equal(p, q) for p, q in zip(x, y)), k
        else:
            assert np.array_equal(x, y), k


@pytest.mark.parametrize("seed,npts,nv", [(3, 20000, 1500), (11, 60000, 5000)])
def test_device_pose_sums_and_argmax_ownership_equal_host_forms(seed, npts, nv):
    from fc_planner_b200 import launch_params, scenes
    pts, nrm, prims = scenes.make_plant(n_points=npts, scale=0.5, seed=seed)
    prm = launch_params("synthetic", seed=seed, resolution=0.2)
    vps = scenes.plant_viewpoints(pts, nrm, prims, nv, prm.viewpoints_distance, seed=seed)
    new = _run(_manager(prm), pts, nrm, vps)
    old = _run(_manager(prm, FCVM_HOST_SUMS=1, FCVM_OWN_SWEEP=1), pts, nrm, vps)
    _same(new, old)
    mixed = _run(_manager(prm, FCVM_HOST_SUMS=1), pts, nrm, vps)
    _same(new, mixed)
    assert len(new["final"][0]) > 10 and (new["pose_init"][:, 3:] != 0).any()


def test_device_pose_sums_match_oracle(oracle_mod, small_plant):
    """setInitViewpoints alone: every averaged pose (FP64) and the ownership state equal the CPU oracle's, and the timers
    say that the device pipeline ran with roughly one ray in four sent to the host's libm."""
    sp = small_plant
    from fc_planner_b200.manager import ViewpointManager
    vm, orc = ViewpointManager(sp["params"]), oracle_mod.Oracle(sp["params"])
    for m in (vm, orc):
        m.setMapPointCloud(sp["points"]); m.setModel(sp["points"]); m.setNormals(sp["points"].astype(np.float64), sp["normals"])
        m.setInitViewpoints(sp["viewpoints"])
    assert np.array_equal(vm.vps_pose(), orc.vps_pose())
    gv, ov = vm.vps_voxcount(), orc.vps_voxcount()
    n = min(len(gv), len(ov))
    assert np.array_equal(gv[:n], ov[:n])
    assert np.array_equal(vm.vps_contri()[:n], orc.vps_contri()[:n])
    assert all(np.array_equal(a, b) for a, b in zip(vm.cover_state(), orc.cover_state()))
    t = vm.timers()
    assert t["pose_sum_pipeline"] > 0 and t["k_eval_E1"] > 0 and t["k_eval_E2"] > 0


def test_pose_sums_pipeline_sub_batches(oracle_mod):
    """Several E1 sub-batches (the pose-sum pipeline of one overlaps the evaluation of the next) give the same poses as one."""
    from fc_planner_b200 import launch_params, scenes
    pts, nrm, prims = scenes.make_plant(n_points=30000, scale=0.5, seed=5)
    prm = launch_params("synthetic", seed=5, resolution=0.2)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 9000, prm.viewpoints_distance, seed=5)     # > 6144: at least two sub-batches
    out = []
    for extra in ({}, {"FCVM_HOST_SUMS": 1}):
        vm = _manager(prm, **extra)
        vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm)
        vm.setInitViewpoints(vps)
        out.append((vm.vps_pose(), vm.vps_voxcount(), vm.vps_contri()))
    assert all(np.array_equal(a, b) for a, b in zip(*out))
