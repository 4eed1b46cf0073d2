"""GPU parity of the pin-hole coverage evaluation (fcvm_cov_*; hcplanner.cpp:1032-1211) against the CPU oracle and
the committed fixtures (tests/golden/cov_<scene>.npz, which record the oracle on the reference's shipped trajectory
poses and its agreement with the reference's own shipped output).

Bar: bit-exact - visible ids per pose, the order-dependent `visible_group` ids, coverage state and rate, and the
algorithmic counters (projections, out-of-image projections)."""
import os
import zlib

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from fc_planner_b200 import Camera, launch_camera, launch_params

pytestmark = pytest.mark.gpu


def _pair(oracle_mod, prm, cam, cloud):
    from fc_planner_b200.coverage import CoverageEvaluator
    ce, orc = CoverageEvaluator(prm, cam), oracle_mod.CoverageOracle(prm, cam)
    ce.setCloud(cloud)
    orc.setCloud(cloud)
    return ce, orc


def _check_both(ce, orc, poses):
    c0 = ce.counters()
    offs, ids, covered = ce.PinHoleCamera(poses)
    c1 = ce.counters() - c0
    o_offs, o_ids, o_cov, o_ctr = orc.pinhole(poses, threads=4)
    assert np.array_equal(offs, o_offs) and np.array_equal(ids, o_ids) and np.array_equal(covered, o_cov)
    assert c1[:3].tolist() == o_ctr.tolist()
    rate, g_offs, g_ids, ever = ce.CoverageEvaluation(poses)
    o_rate, og_offs, og_ids, o_ever = orc.CoverageEvaluation(poses)
    assert rate == o_rate and np.array_equal(g_offs, og_offs) and np.array_equal(g_ids, og_ids) and np.array_equal(ever, o_ever)
    cov_rate, cov2 = ce.path_coverage(poses)
    assert np.array_equal(cov2, o_cov) and cov_rate == float(o_cov.sum()) / len(o_cov)
    return len(ids), len(g_ids), ce.counters() - c0


@pytest.mark.parametrize("scene", ["mbs", "pipe", "christ"])
def test_shipped_trajectories_bit_exact(oracle_mod, scene):
    from fc_planner_b200.coverage import CoverageEvaluator
    _, zs = load_golden(scene)
    z = np.load(os.path.join(GOLDEN, f"cov_{scene}.npz"))
    ce = CoverageEvaluator(launch_params(scene), launch_camera(scene))
    ce.setCloud(zs["map_cloud"])
    rate, offs, ids, ever = ce.CoverageEvaluation(z["poses"])
    assert rate == float(z["rate"])
    assert np.array_equal(offs, z["group_offsets"]) and np.array_equal(ids, z["group_ids"])
    assert np.array_equal(np.packbits(ever, bitorder="little"), z["ever"])
    p_offs, p_ids, covered = ce.PinHoleCamera(z["poses"])
    assert np.array_equal(p_offs, z["pin_offsets"]) and zlib.crc32(p_ids.tobytes()) == int(z["pin_ids_crc"])
    assert np.array_equal(np.packbits(covered, bitorder="little"), z["covered"])
    ctr = ce.counters()
    assert ctr[0] == 2 * len(z["poses"]) * len(zs["map_cloud"])
    assert ctr[1] == 2 * z["counters"][1] and ctr[2] == 2 * z["counters"][2]
    cp = ce.chunk_points()
    assert ctr[4] < 0.5 * 2 * len(z["poses"]) * ((len(zs["map_cloud"]) + cp - 1) // cp)    # the chunk cull does cull


def test_dense_collisions_small_image(oracle_mod, small_plant):
    """A 32 x 24 image: almost every pixel is contested, so the in-chunk ordering rule, the displaced marks and the
    cross-pose cover-state sweep all carry weight."""
    prm = small_plant["params"]
    cam = Camera(fx=417.0467, fy=461.0951, cx=16.0, cy=12.0, fov_h=prm.fov_h, fov_w=prm.fov_w)
    ce, orc = _pair(oracle_mod, prm, cam, small_plant["points"])
    from fc_planner_b200 import scenes
    poses = scenes.poses_from_viewpoints(small_plant["viewpoints"])
    rng = np.random.default_rng(5)
    poses = np.concatenate([poses, poses[rng.permutation(len(poses))[:200]] + rng.normal(0, 0.05, (200, 5))])
    n_vis, n_grp, ctr = _check_both(ce, orc, poses)
    assert n_vis > 20_000 and n_grp > 1000
    assert ctr[5] > 100                       # duplicate-pixel slow paths were exercised


def test_ties_and_duplicates(oracle_mod):
    """Coincident points (equal distances on one pixel): the earliest index keeps the pixel, in a chunk and across chunks."""
    prm, cam = launch_params("mbs"), launch_camera("mbs")
    rng = np.random.default_rng(11)
    base = rng.uniform(-4, 4, (700, 3)).astype(np.float32)
    cloud = np.concatenate([base, base[::-1], base[rng.permutation(700)], base + np.float32(1e-4)])    # 2800 points: 3 chunks
    ce, orc = _pair(oracle_mod, prm, cam, cloud)
    poses = np.concatenate([rng.uniform(-9, 9, (150, 3)), rng.uniform(-1.2, 1.2, (150, 1)), rng.uniform(-np.pi, np.pi, (150, 1))], axis=1)
    poses[:, 3:] = np.where(rng.random((150, 1)) < 0.5, poses[:, 3:], 0.0)
    n_vis, n_grp, _ = _check_both(ce, orc, poses)
    assert n_vis > 5000


@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 4097])
def test_ragged_cloud_sizes(oracle_mod, n):
    prm, cam = launch_params("mbs"), launch_camera("mbs")
    rng = np.random.default_rng(n)
    cloud = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    ce, orc = _pair(oracle_mod, prm, cam, cloud)
    poses = np.concatenate([rng.uniform(-8, 8, (40, 3)), rng.uniform(-1.0, 1.0, (40, 1)), rng.uniform(-np.pi, np.pi, (40, 1))], axis=1)
    _check_both(ce, orc, poses)
    _check_both(ce, orc, poses[:1])


def test_large_synthetic_cloud(oracle_mod):
    """200 k-point plant, 0.1 m-class density: many chunks, culling on, a few hundred passing points per pose."""
    from fc_planner_b200 import scenes
    pts, nrm, prims = scenes.make_plant(n_points=200_000, scale=0.5, seed=2)
    prm, cam = launch_params("mbs", z_ground=0), launch_camera("mbs")
    vps = scenes.plant_viewpoints(pts, nrm, prims, 600, prm.viewpoints_distance, seed=2)
    ce, orc = _pair(oracle_mod, prm, cam, pts)
    n_vis, n_grp, ctr = _check_both(ce, orc, scenes.poses_from_viewpoints(vps))
    assert n_vis > 100_000


def test_no_cloud_is_an_error():
    from fc_planner_b200 import _lib
    from fc_planner_b200.coverage import CoverageEvaluator
    ce = CoverageEvaluator(launch_params("mbs"), launch_camera("mbs"))
    with pytest.raises(_lib.FcvmError) as e:
        ce.PinHoleCamera(np.zeros((1, 5)))
    assert e.value.code == -2
