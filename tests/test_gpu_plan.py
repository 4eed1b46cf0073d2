"""GPU parity of the planner-grid loops (fcvm_grid_*) against the CPU oracle (oracle/fcvm_oracle_plan.hpp, pinned to the
reference's RayCaster in tests/test_oracle_plan.py).  Bar: bit-exact verdicts, crossing counts, nearest ids, FP64
distances; voxel-lookup counts equal for the two ray loops."""
import numpy as np
import pytest

from test_oracle_plan import _grid, _segments

pytestmark = pytest.mark.gpu


def _pair(oracle_mod, *a):
    from fc_planner_b200.planner_grid import PlannerGrid
    return PlannerGrid(*a), oracle_mod.PlanOracle(*a)


@pytest.mark.parametrize("seed,dims,res", [(0, (40, 36, 30), 0.4), (1, (33, 61, 18), 0.25), (2, (5, 7, 3), 1.0)])
def test_line_of_sight_all_pairs(oracle_mod, seed, dims, res):
    origin, res, dims, hc, inflate, internal = _grid(seed, dims, res)
    pg, orc = _pair(oracle_mod, origin, res, dims, hc, inflate, internal)
    rng = np.random.default_rng(seed + 10)
    a, b = _segments(rng, origin, res, dims, 150)
    P = np.concatenate([a, b])
    safe, dist = pg.lineOfSight(P)
    o_safe, o_dist = orc.lineOfSight(P, threads=4)
    assert np.array_equal(safe, o_safe) and np.array_equal(dist, o_dist)
    assert pg.counters()[0] == orc.counters()[0] and pg.counters()[1] == orc.counters()[1]
    assert 0 < safe.sum() < safe.size


def test_line_of_sight_structured_scene(oracle_mod):
    """Mostly empty grid with a few solid blocks and thin walls: the rays that the free-box certificate ends (before the
    march and in flight), blocked rays, lattice-aligned and out-of-map positions, all against the literal oracle."""
    dims = np.array([96, 80, 64], np.int32)
    res, origin = 0.4, np.array([-3.1, -2.7, -1.3])
    g = np.zeros(tuple(dims), np.int8)
    g[30:40, 20:60, 10:50] = 1; g[60:62, :, 5:40] = 1; g[10:20, 65:70, :] = 1; g[70:90, 30:34, 30:34] = 1; g[5, 5, 5] = 1
    inflate = g.reshape(-1).copy()
    internal = np.zeros_like(inflate); internal.reshape(tuple(dims))[45:50, 40:45, 50:60] = 1
    pg, orc = _pair(oracle_mod, origin, res, dims, None, inflate, internal)
    rng = np.random.default_rng(7)
    lo, hi = origin, origin + res * dims
    P = rng.uniform(lo + 0.5, hi - 0.5, (640, 3))
    P[:40] = np.round(P[:40] / res) * res                     # on lattice planes: the literal marcher
    P[40:60] = rng.uniform(lo - 2.0, hi + 2.0, (20, 3))       # some outside the map
    P[60:80] = (np.floor(P[60:80] / res) + 0.5) * res         # voxel centres
    P[80:90, 0] = P[90:100, 0]                                # axis-aligned pairs
    safe, dist = pg.lineOfSight(P)
    o_safe, o_dist = orc.lineOfSight(P, threads=8)
    assert np.array_equal(safe, o_safe) and np.array_equal(dist, o_dist)
    assert pg.counters()[0] == orc.counters()[0] and pg.counters()[1] == orc.counters()[1]
    assert 0.2 < safe.mean() < 0.9
    # a second call on the same grid (table cached) and a call after the planes changed (table rebuilt)
    safe2, _ = pg.lineOfSight(P[:100], want_dist=False)
    assert np.array_equal(safe2, o_safe[:100, :100])
    outers = np.array([[15.9 + 0.1 * k, 14.3, 17.0, 0.0, 0.0, 1.0] for k in range(8)])       # rays up through the internal block clear its marks
    before = pg.buffers()[2].sum()
    pg.OuterCheck(outers, 8.0); orc.OuterCheck(outers, 8.0)
    assert pg.buffers()[2].sum() < before and np.array_equal(pg.buffers()[2], orc.buffers()[2])
    Q = np.concatenate([P[100:300], np.array([[15.9, 14.3, 17.0], [15.9, 14.3, 24.0], [16.3, 14.3, 17.5], [16.3, 14.3, 23.5]])])
    safe3, _ = pg.lineOfSight(Q, want_dist=False)
    o_safe3 = orc.lineOfSight(Q, threads=8)[0]
    assert np.array_equal(safe3, o_safe3) and o_safe3[200, 201] == 1


def test_light_state_and_safety_box(oracle_mod):
    origin, res, dims, hc, inflate, internal = _grid(3, (48, 40, 36), 0.4, fill=0.05)
    pg, orc = _pair(oracle_mod, origin, res, dims, hc, inflate, internal)
    rng = np.random.default_rng(4)
    a, b = _segments(rng, origin, res, dims, 2000)
    rosa = b[:300].astype(np.float32)
    rosa[10] = rosa[11]                                   # a tie in the nearest-vertex search: lowest index wins on both sides
    det, cross, near = pg.lightStateDet(a, rosa)
    o_det, o_cross, o_near = orc.lightStateDet(a, rosa)
    assert np.array_equal(near, o_near) and np.array_equal(cross, o_cross) and np.array_equal(det, o_det)
    assert pg.counters()[0] == orc.counters()[0]
    assert 0 < det.sum() < len(det)
    vps = np.concatenate([a, b]).astype(np.float32)
    for sr in (1.0, 1.5, 0.3):
        ok = pg.safetyBox(vps, sr)
        assert np.array_equal(ok, orc.safetyBox(vps, sr))
        assert 0 < ok.sum() < len(ok)


def test_empty_buffers_and_single_point(oracle_mod):
    pg, orc = _pair(oracle_mod, [0, 0, 0], 0.5, [12, 12, 12], None, None, None)
    P = np.array([[1.0, 1.0, 1.0]])
    safe, dist = pg.lineOfSight(P)
    assert safe.tolist() == [[1]] and dist.tolist() == [[0.0]]
    det, cross, near = pg.lightStateDet(P, np.zeros((0, 3), np.float32))
    assert det.tolist() == [0] and near.tolist() == [-1]
