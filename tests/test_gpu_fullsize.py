"""Full-size checks at BASELINE.json's synthetic configuration (configs[4] geometry: P = 1 000 000
surface points, 0.1 m grid, 100 MB bit-packed): size-independent properties of the CUDA path plus
an oracle spot check on a few viewpoints (the oracle needs ~0.2 s per viewpoint at this size).
Bar: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    from fc_planner_b200 import launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
    prm = launch_params("synthetic", seed=0, resolution=0.1)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 768, prm.viewpoints_distance, seed=42)
    poses = scenes.poses_from_viewpoints(vps)
    vm = ViewpointManager(prm)
    vm.setMapPointCloud(pts); vm.setModel(pts)
    return dict(pts=pts, nrm=nrm, prm=prm, poses=poses, vm=vm)


def _popc(rows):
    return np.unpackbits(rows.view(np.uint8), axis=1).sum(axis=1, dtype=np.int64)


def test_counts_rows_csr_agree_and_are_deterministic(full):
    vm, poses = full["vm"], full["poses"]
    cnt, rows = vm.evaluate(1, poses)
    assert len(full["pts"]) == 1_000_000 and rows.shape == (len(poses), vm.words())
    assert np.array_equal(cnt, _popc(rows))
    assert cnt.sum() > 1_000_000                                    # not vacuous: > 1300 visible points per viewpoint
    assert not rows[:, (len(full["pts"]) + 31) // 32:].any()        # padding words stay zero
    cnt2, rows2 = vm.evaluate(1, poses)
    assert np.array_equal(rows, rows2) and np.array_equal(cnt, cnt2)          # idempotent / deterministic (atomicOr order-free)
    c3, offs, idx = vm.evaluate_csr(1, poses)
    assert np.array_equal(c3, cnt) and offs[-1] == cnt.sum()
    for i in (0, 17, len(poses) - 1):
        want = np.flatnonzero(np.unpackbits(rows[i].view(np.uint8), bitorder="little"))
        assert np.array_equal(idx[offs[i]:offs[i + 1]], want)       # ascending point indices
    full["rows1"], full["cnt1"] = rows, cnt


def test_result_is_independent_of_batching(full):
    """A viewpoint's row does not depend on which batch (shard) it is evaluated in."""
    vm, poses = full["vm"], full["poses"]
    rows = full["rows1"]
    for lo, hi in ((0, 1), (5, 133), (500, 768)):
        _, part = vm.evaluate(1, poses[lo:hi])
        assert np.array_equal(part, rows[lo:hi])
    perm = np.random.default_rng(0).permutation(len(poses))[:200]
    _, part = vm.evaluate(1, poses[perm])
    assert np.array_equal(part, rows[perm])


def test_stage_relations(full):
    vm, poses = full["vm"], full["poses"][:256]
    rows1 = full["rows1"][:256]
    _, rows3 = vm.evaluate(3, poses)
    assert not (rows1 & ~rows3).any()          # E1 = E3 plus the range re-check at the same pose: E1 is a subset of E3
    poses2 = poses.copy(); poses2[:, 3:] += 0.03
    cnt2, rows2 = vm.evaluate(2, poses2, restrict_rows=rows1)
    assert not (rows2 & ~rows1).any()          # E2 only re-tests E1's set
    assert 0 < cnt2.sum() < full["cnt1"][:256].sum()
    _, rows4 = vm.evaluate(4, poses)
    assert rows4.any()


def test_oracle_spot_check_at_full_size(full, oracle_mod):
    orc = oracle_mod.Oracle(full["prm"])
    orc.setMapPointCloud(full["pts"]); orc.setModel(full["pts"])
    sel = np.array([0, 101, 333, 600, 767])
    poses = full["poses"][sel]
    vm = full["vm"]
    W = (len(full["pts"]) + 31) // 32
    for stage in (1, 4):
        c0 = vm.counters()
        cnt, rows = vm.evaluate(stage, poses)
        c1 = vm.counters() - c0
        ocnt, orows, octr = orc.evaluate(stage, poses, threads=5)
        assert np.array_equal(rows[:, :W], orows) and np.array_equal(cnt, ocnt)
        assert np.array_equal(c1[:5], octr[:5]) and c1[4] == 0       # pair tests, rays, lookups, walk lookups; no cap trips
    assert cnt.sum() > 1000


def test_ownership_sweep_properties(full):
    """After the ownership sweep over a batch (viewpoint_manager.cpp:500-528): every seen point has an
    owner, an owner's vox count is the number of points it owns, cover_contrib is the largest
    contribute_voxel_num among the point's seers, ties go to the earlier viewpoint."""
    vm = full["vm"]
    poses = full["poses"][:128]
    vm.reset()
    vm.setMapPointCloud(full["pts"]); vm.setModel(full["pts"])
    vm.stage_poses(poses)
    vm.evaluate_resident(4, sync=True)
    ids = np.arange(len(poses), dtype=np.int32)
    vm.own_resident(ids)
    cnt, rows = vm.evaluate(4, poses)
    cov, contrib, owner = vm.cover_state()
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :len(cov)].astype(bool)
    seen = bits.any(axis=0)
    assert np.array_equal(cov.astype(bool), seen)
    best = np.where(bits, cnt[:, None].astype(np.int32), np.int32(-1))
    assert np.array_equal(contrib[seen], best.max(axis=0)[seen])
    assert np.array_equal(owner[seen], best.argmax(axis=0)[seen])     # argmax = first maximum = earliest viewpoint
    vox = vm.vps_voxcount()
    assert np.array_equal(vox[:len(poses)], np.bincount(owner[seen], minlength=len(poses)))
    assert vox.sum() == seen.sum()


def test_csr_streamed_path_and_caller_buffers(full):
    """fcvm_evaluate_csr on a batch large enough for the sub-batched path (index lists extracted and copied out while the
    next sub-batch is evaluated): every list equals the bits of the row; preallocated page-locked output buffers receive
    the lists directly; a buffer that is too small keeps a valid prefix and the call reports the full size."""
    from fc_planner_b200.manager import pinned_empty
    vm = full["vm"]
    poses = np.concatenate([full["poses"]] * 3)               # 2304 rows: the sub-batched path needs >= 2 x 1024
    n = len(poses)
    assert n >= 2048
    rows, cnt = np.concatenate([full["rows1"]] * 3), np.concatenate([full["cnt1"]] * 3)
    total = int(cnt.sum())
    c, offs, idx = vm.evaluate_csr(1, poses)
    assert np.array_equal(c, cnt) and len(idx) == total and np.array_equal(np.diff(offs), cnt)
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")
    r, j = np.nonzero(bits)                                   # row-major: ascending point index inside each row
    assert np.array_equal(idx, j.astype(np.uint32))
    out = (pinned_empty((n,), np.int32), pinned_empty((n + 1,), np.int64), pinned_empty((total + 5,), np.uint32))
    c2, offs2, idx2 = vm.evaluate_csr(1, poses, out=out)
    assert np.array_equal(c2, cnt) and np.array_equal(offs2, offs) and np.array_equal(idx2, idx)
    small = (np.zeros(n, np.int32), np.zeros(n + 1, np.int64), np.full(total // 3, 0xFFFFFFFF, np.uint32))
    with pytest.raises(ValueError):
        vm.evaluate_csr(1, poses, out=small)
    assert np.array_equal(small[2], idx[:total // 3]) and np.array_equal(small[1], offs)
