"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Bar: bit-exact.  Masks, counts, ownership state and the selected viewpoint set are integers and
must be identical; poses are FP64 produced by the same libm on the host from identical masks, so
they must be identical bit for bit as well (tolerance = 0).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(small_plant, oracle_mod, **over):
    from fc_planner_b200.manager import ViewpointManager
    prm = small_plant["params"].copy(**over)
    vm = ViewpointManager(prm)
    orc = oracle_mod.Oracle(prm)
    for m in (vm, orc):
        m.setMapPointCloud(small_plant["points"])
        m.setModel(small_plant["points"])
        m.setNormals(small_plant["points"].astype(np.float64), small_plant["normals"])
    return vm, orc, prm


def _perturbed_poses(small_plant, n=300, seed=3):
    from fc_planner_b200 import scenes
    poses = scenes.poses_from_viewpoints(small_plant["viewpoints"][:n])
    rng = np.random.default_rng(seed)
    poses[:, :3] += rng.normal(0, 0.3, (len(poses), 3))        # arbitrary doubles, not float-rounded
    poses[:, 3:] += rng.normal(0, 0.2, (len(poses), 2))
    return poses


@pytest.mark.parametrize("stage", [1, 3, 4])
def test_evaluator_masks_bit_exact(small_plant, oracle_mod, stage):
    vm, orc, _ = _mk(small_plant, oracle_mod)
    poses = _perturbed_poses(small_plant)
    c0 = vm.counters()
    cnt, rows = vm.evaluate(stage, poses)
    c1 = vm.counters() - c0
    ocnt, orows, octr = orc.evaluate(stage, poses)
    W = orows.shape[1]
    assert np.array_equal(rows[:, :W], orows)
    assert not rows[:, W:].any()
    assert np.array_equal(cnt, ocnt)
    assert cnt.sum() > 1000                       # the case is not vacuous
    # algorithmic counters: pair tests, rays, voxel lookups, walk lookups identical; no cap trips
    assert c1[0] == octr[0] and c1[1] == octr[1] and c1[2] == octr[2] and c1[3] == octr[3]
    assert c1[4] == 0 and octr[4] == 0


def test_e1_visible_range_band(small_plant, oracle_mod):
    """E1's `ray_length <= visible_range` is decided from the squared length outside a 1e-12 band around visible_range^2 and by
    the literal sqrt inside it: viewpoints placed at visible_range (1 + k 2e-14) from a point they see, k = -20 .. 20."""
    vr = 6.0
    vm, orc, prm = _mk(small_plant, oracle_mod, visible_range=vr)
    base = _perturbed_poses(small_plant, n=120, seed=5)
    _, rows0 = vm.evaluate(1, base)                       # with the short range: who sees what, roughly
    wide_vm, _, _ = _mk(small_plant, oracle_mod)          # default range (= max_dist): the visible sets to pick targets from
    cnt_w, rows_w = wide_vm.evaluate(1, base)
    pts = small_plant["points"].astype(np.float64)
    poses, in_band = [], 0
    for i in np.flatnonzero(cnt_w > 0)[:40]:
        seen = np.flatnonzero(np.unpackbits(rows_w[i].view(np.uint8), bitorder="little"))
        j = int(seen[len(seen) // 2])
        d = pts[j] - base[i, :3]
        u = d / np.linalg.norm(d)
        for k in range(-20, 21):
            p = base[i].copy()
            p[:3] = pts[j] - u * (vr * (1.0 + k * 2e-14))
            poses.append(p)
            dd = pts[j] - p[:3]
            n2 = (dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2]
            in_band += abs(n2 / (vr * vr) - 1.0) < 1e-12
    poses = np.array(poses)
    assert in_band > 200                                   # the literal path is exercised
    cnt, rows = vm.evaluate(1, poses)
    ocnt, orows, _ = orc.evaluate(1, poses)
    assert np.array_equal(rows[:, :orows.shape[1]], orows) and np.array_equal(cnt, ocnt)
    assert 0 < cnt.sum()


@pytest.mark.parametrize("n_points,order", [(4999, "shuffled"), (1023, "shuffled"), (4133, "sorted_z"), (33, "as_is")])
def test_culls_exact_for_any_point_order(small_plant, oracle_mod, n_points, order):
    """The tile- and group-level box culls (k_eval: one viewpoint / one 32-point group per lane) must drop nothing the
    reference would test positively, whatever the order of the model cloud and however ragged the last tile and group
    are: shuffled clouds (every box spans the scene: nothing culled), z-sorted clouds (thin slabs: most culled), sizes
    that leave partial groups and a single partial tile.  Masks, counts and the algorithmic counters against the oracle."""
    from fc_planner_b200.manager import ViewpointManager
    rng = np.random.default_rng(n_points)
    pts = small_plant["points"][rng.choice(len(small_plant["points"]), n_points, replace=False)]
    if order == "sorted_z":
        pts = pts[np.argsort(pts[:, 2], kind="stable")]
    elif order == "as_is":
        pts = np.sort(pts, axis=0)[:n_points]
    pts = np.ascontiguousarray(pts)
    prm = small_plant["params"]
    vm, orc = ViewpointManager(prm), oracle_mod.Oracle(prm)
    for m in (vm, orc):
        m.setMapPointCloud(small_plant["points"])        # the full plant occludes
        m.setModel(pts)
    poses = _perturbed_poses(small_plant, n=200, seed=n_points)
    W = (n_points + 31) // 32
    rows1 = None
    for stage in (1, 2, 3, 4):
        c0 = vm.counters()
        if stage == 2:
            cnt, rows = vm.evaluate(2, poses, restrict_rows=rows1)
            ocnt, orows, octr = orc.evaluate(2, poses, restrict_rows=rows1[:, :W])
        else:
            cnt, rows = vm.evaluate(stage, poses)
            ocnt, orows, octr = orc.evaluate(stage, poses)
        c1 = vm.counters() - c0
        assert np.array_equal(rows[:, :W], orows) and not rows[:, W:].any(), stage
        assert np.array_equal(cnt, ocnt), stage
        assert np.array_equal(c1[:5], octr[:5]), (stage, c1[:5], octr[:5])
        if stage == 1:
            rows1 = rows
            assert cnt.sum() > 50


def test_e2_restricted_to_e1_rows(small_plant, oracle_mod):
    vm, orc, _ = _mk(small_plant, oracle_mod)
    poses = _perturbed_poses(small_plant)
    _, rows1 = vm.evaluate(1, poses)
    poses2 = poses.copy()
    poses2[:, 3:] += 0.05                          # E2 runs with the averaged pose, i.e. a different FOV
    c0 = vm.counters()
    cnt2, rows2 = vm.evaluate(2, poses2, restrict_rows=rows1)
    c1 = vm.counters() - c0
    W = (orc.P + 31) // 32
    ocnt2, orows2, octr = orc.evaluate(2, poses2, restrict_rows=rows1[:, :W])
    assert np.array_equal(rows2[:, :W], orows2)
    assert np.array_equal(cnt2, ocnt2)
    assert (rows2 & ~rows1).sum() == 0            # E2 is a subset of E1
    assert np.array_equal(c1[:5], octr[:5])


def _compare_state(vm, orc):
    assert np.array_equal(vm.vps_pose(), orc.vps_pose())                    # FP64, bit-exact
    ov = orc.vps_voxcount(); gv = vm.vps_voxcount()
    n = min(len(ov), len(gv))
    assert np.array_equal(gv[:n], ov[:n]) and not gv[n:].any() and not ov[n:].any()
    assert np.array_equal(vm.vps_contri(), orc.vps_contri())
    gc, oc = vm.cover_state(), orc.cover_state()
    assert np.array_equal(gc[0], oc[0])
    assert np.array_equal(gc[1], oc[1])
    assert np.array_equal(gc[2], oc[2])


def _trace_dict(tr, W):
    d = {}
    for st, vid, pose, row in tr:
        d.setdefault((st, vid), []).append((pose.copy(), row[:W].copy()))
    return d


@pytest.mark.parametrize("pose_update", [1, 0])
def test_set_init_viewpoints(small_plant, oracle_mod, pose_update):
    vm, orc, _ = _mk(small_plant, oracle_mod, pose_update=pose_update)
    vm.trace_enable(); orc.trace_enable()
    vm.setInitViewpoints(small_plant["viewpoints"]); orc.setInitViewpoints(small_plant["viewpoints"])
    _compare_state(vm, orc)
    W = (orc.P + 31) // 32
    gt, ot = _trace_dict(vm.trace(), W), _trace_dict(orc.trace(), W)
    assert gt.keys() == ot.keys()
    for k in ot:
        assert len(gt[k]) == len(ot[k])
        for (gp, gr), (op, orow) in zip(gt[k], ot[k]):
            assert np.array_equal(gp, op), k
            assert np.array_equal(gr, orow), k


@pytest.mark.parametrize("init", ["explicit", "from_normals"])
def test_update_viewpoints_full_sequence(small_plant, oracle_mod, init):
    """reset -> setMapPointCloud -> setModel -> setNormals -> [setInitViewpoints] -> updateViewpoints
    -> getUpdatedViewpoints -> getUncoveredModelCloud (hcplanner.cpp:709-716)."""
    vm, orc, _ = _mk(small_plant, oracle_mod)
    vm.trace_enable(); orc.trace_enable()
    if init == "explicit":
        vm.setInitViewpoints(small_plant["viewpoints"]); orc.setInitViewpoints(small_plant["viewpoints"])
    assert vm.updateViewpoints() == 0 and orc.updateViewpoints() == 0
    _compare_state(vm, orc)
    gi, gp, gx = vm.final_arrays()
    oi, op, ox = orc.getUpdatedViewpoints()
    assert len(oi) > 5
    assert np.array_equal(gi, oi)            # the selected viewpoint set, in the reference's order
    assert np.array_equal(gp, op)            # poses bit-exact
    assert np.array_equal(gx, ox)
    assert np.array_equal(vm.getUncoveredModelCloud(), orc.getUncoveredModelCloud())
    W = (orc.P + 31) // 32
    gt, ot = _trace_dict(vm.trace(), W), _trace_dict(orc.trace(), W)
    assert gt.keys() == ot.keys()
    bad = 0
    for k in ot:
        assert len(gt[k]) == len(ot[k]), k
        for (gpz, gr), (opz, orow) in zip(gt[k], ot[k]):
            bad += int(not np.array_equal(gr, orow)) + int(not np.array_equal(gpz, opz))
    assert bad == 0
    gc, oc = vm.counters(), orc.counters()
    assert gc[4] == 0 and oc[4] == 0          # no marcher cap trips on either side
    assert gc[1] == oc[1] and gc[2] == oc[2]  # same rays, same voxel lookups
    assert gc[7] == oc[7]                      # same safety_check lookups


def test_second_update_after_reset_matches(small_plant, oracle_mod):
    """reset() keeps the prune maps' bucket arrays (clear(), not reconstruction): a second run on the
    same handle must still match the oracle handle that went through the same history."""
    vm, orc, _ = _mk(small_plant, oracle_mod)
    for m in (vm, orc):
        m.updateViewpoints()
        m.reset()
        m.setMapPointCloud(small_plant["points"]); m.setModel(small_plant["points"])
        m.setNormals(small_plant["points"].astype(np.float64), small_plant["normals"])
        m.setInitViewpoints(small_plant["viewpoints"][:150])
        m.updateViewpoints()
    gi, gp, gx = vm.final_arrays()
    oi, op, ox = orc.getUpdatedViewpoints()
    assert np.array_equal(gi, oi) and np.array_equal(gp, op) and np.array_equal(gx, ox)


def test_error_convention(small_plant):
    from fc_planner_b200.manager import ViewpointManager
    vm = ViewpointManager(small_plant["params"])
    assert vm.setModel(np.zeros((0, 3), np.float32)) == -1          # "Input model is empty!!" -> log + return
    assert vm.updateViewpoints() == -2                               # TaskState not ready
    assert len(vm.getUpdatedViewpoints()) == 0


@pytest.mark.parametrize("scene", ["mbs", "pipe", "christ", "pipe4x"])
def test_shipped_scene_matches_golden(scene):
    """BASELINE.json configs[0..2]: the full 8-call sequence on the reference's own MBS / pipe /
    Christ point clouds against the committed golden fixtures (tests/golden/*.npz, generated by the
    oracle from /root/reference's .pcd files): every evaluator mask, pose, count, the ownership
    state, the pruned viewpoint set in the reference's order, and the uncovered cloud.  Bit-exact."""
    from conftest import compare_with_golden, load_golden, run_sequence
    from fc_planner_b200.manager import ViewpointManager
    prm, z = load_golden(scene)
    vm = ViewpointManager(prm)
    run_sequence(vm, z)
    compare_with_golden(vm, z, vm.final_arrays())
    assert vm.counters()[7] == z["counters"][7]       # same safety_check lookups on the host


def test_lattice_aligned_viewpoints(small_plant, oracle_mod):
    """Viewpoints exactly on voxel-lattice corners: the inputs on which the reference's marcher can
    run away (no termination; tests/test_oracle_vs_ref.py).  Oracle and kernel cap such rays and
    count them (counter 4); every viewpoint without a capped ray must still match bit for bit, and
    nothing may fault."""
    from fc_planner_b200 import scenes
    vm, orc, prm = _mk(small_plant, oracle_mod)
    poses = scenes.poses_from_viewpoints(small_plant["viewpoints"][:160])
    poses[:, :3] = np.round(poses[:, :3] / prm.resolution) * prm.resolution
    c0 = vm.counters()
    cnt, rows = vm.evaluate(1, poses)
    gpu_trips = int((vm.counters() - c0)[4])
    W = (orc.P + 31) // 32
    clean, capped = 0, 0
    for i in range(len(poses)):
        ocnt, orows, octr = orc.evaluate(1, poses[i:i + 1])
        if octr[4] == 0:
            assert np.array_equal(rows[i, :W], orows[0]) and cnt[i] == ocnt[0], i
            clean += 1
        else:
            capped += 1
    assert clean > 100
    assert (gpu_trips > 0) == (capped > 0) or capped == 0
    print(f"lattice-aligned: {clean} clean viewpoints bit-exact, {capped} with capped rays (oracle), gpu cap trips {gpu_trips}")


def test_offset_walk_with_lattice_aligned_and_border_points(oracle_mod):
    """E2 / E4's surface-offset walk starts at the model point: points exactly on voxel-lattice planes (the fast walk must not
    take them: the literal marcher decides), points in general position, and viewpoints outside the map, on a 0.5 m grid
    where the lattice is exact in float32.  Masks, counts and the walk's look-up counter against the oracle."""
    from fc_planner_b200 import launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    pts, nrm, prims = scenes.make_plant(n_points=4000, scale=0.4, seed=2)
    pts = pts.copy()
    pts[:1200] = np.round(pts[:1200] / 0.5) * 0.5                # on lattice corners / edges / faces
    pts[1200:1800, 0] = np.round(pts[1200:1800, 0] / 0.5) * 0.5  # on x planes only
    prm = launch_params("mbs", z_ground=0, seed=7, resolution=0.5)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 200, prm.viewpoints_distance, seed=1)
    poses = scenes.poses_from_viewpoints(vps)
    rng = np.random.default_rng(9)
    poses[:, :3] += rng.normal(0, 0.3, (len(poses), 3))
    poses[:8, :3] += 60.0                                        # outside the map: the walk's end voxel is not in the grid
    vm, orc = ViewpointManager(prm), oracle_mod.Oracle(prm)
    for m in (vm, orc):
        m.setMapPointCloud(pts); m.setModel(pts); m.setNormals(pts.astype(np.float64), nrm)
    _, e1 = vm.evaluate(1, poses)
    for stage, restrict in ((4, None), (2, e1)):
        c0 = vm.counters()
        cnt, rows = vm.evaluate(stage, poses, restrict_rows=restrict)
        c1 = vm.counters() - c0
        ocnt, orows, octr = orc.evaluate(stage, poses, restrict_rows=None if restrict is None else restrict[:, :(orc.P + 31) // 32])
        W = orows.shape[1]
        ok = octr[4] == 0                                         # no capped march on the oracle's side: everything must match
        assert c1[4] == octr[4]
        if ok:
            assert np.array_equal(rows[:, :W], orows) and np.array_equal(cnt, ocnt), stage
            assert c1[1] == octr[1] and c1[2] == octr[2] and c1[3] == octr[3], stage
        assert cnt.sum() > 500


def test_safety_check_batch_matches_oracle(small_plant, oracle_mod):
    """viewpointSafetyCheck on the device (k_safety): verdicts and the number of safety_check probes (early exit
    included) against the oracle, for positions near, inside, far from and outside the map, with and without ground."""
    for over in (dict(), dict(z_ground=1, ground_pos=-2.0), dict(safe_radius=1.3), dict(safe_radius=0.9)):
        vm, orc, prm = _mk(small_plant, oracle_mod, **over)
        rng = np.random.default_rng(17)
        pts = small_plant["points"]
        lo, hi = pts.min(0), pts.max(0)
        q = np.concatenate([
            small_plant["viewpoints"][:300, :3],                                   # 8 m off the surface
            pts[rng.integers(0, len(pts), 200)] + rng.normal(0, 1.5, (200, 3)),    # close to the surface
            pts[rng.integers(0, len(pts), 100)] + rng.normal(0, 3.5, (100, 3)),    # around safe_radius
            rng.uniform(lo - 30, hi + 30, (200, 3)),                               # up to outside the map
            np.array([[np.nan, 0, 0], [np.inf, 0, 0], [1e30, 1e30, 1e30]]),
        ]).astype(np.float32)
        c0, o0 = vm.counters()[7], orc.counters()[7]
        ok = vm.safety_check(q)
        ook = orc.safety_check(q)
        assert np.array_equal(ok, ook)
        assert vm.counters()[7] - c0 == orc.counters()[7] - o0 > 0
        assert 0 < ok.sum() < len(ok)


def test_device_grid_equals_host_build(small_plant, oracle_mod):
    """initHCMap voxelisation on the device (k_voxelise) against the oracle's grid: origin, voxel counts, every bit."""
    vm, orc, prm = _mk(small_plant, oracle_mod)
    g_org, g_dims, g_bits = vm.grid()
    o_org, o_dims, o_bits = orc.grid()
    assert np.array_equal(g_org, o_org) and np.array_equal(g_dims, o_dims) and np.array_equal(g_bits, o_bits)
    assert g_bits.any()
