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
