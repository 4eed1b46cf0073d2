"""SDFMap::SetOcc / InternalSpace / OuterCheck on the device (fcvm_grid_*; plan_env/src/sdf_map.cpp:263-473) against the CPU
oracle's statement-by-statement restatement (oracle/fcvm_oracle_plan.hpp): the three voxel buffers must be identical byte for
byte, and the loops that read them afterwards (search_Path's straight-line test, lightStateDet) must agree as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu



def _skeleton_scene(seed, res):
    """A synthetic 'ROSA output': a few skeleton segments, each with a tube-like cloud of surface points around it."""
    rng = np.random.default_rng(seed)
    verts = np.array([[0, 0, 2], [0, 0, 14], [9, 3, 18], [-7, 4, 17], [0, 0, 14.0]]) + rng.normal(0, 0.2, (5, 3))
    segs = [(0, 1), (1, 2), (1, 3)]
    clouds, ends = [], []
    for a, b in segs:
        p0, p1 = verts[a], verts[b]
        axis = (p1 - p0) / np.linalg.norm(p1 - p0)
        u = np.cross(axis, [1.0, 0.3, 0.2]); u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        n = 1500
        t = rng.uniform(-0.05, 1.05, n)[:, None]
        th = rng.uniform(0, 2 * np.pi, n)[:, None]
        r = 2.0 + 0.6 * np.sin(3 * th) + rng.normal(0, 0.05, (n, 1))
        clouds.append((p0 + t * (p1 - p0) + r * (np.cos(th) * u + np.sin(th) * v)).astype(np.float32))
        ends.append(np.concatenate([p0, p1]))
    allp = np.concatenate(clouds)
    origin = allp.min(0).astype(np.float64) - 4.0
    dims = np.ceil((allp.max(0) + 4.0 - origin) / res).astype(np.int32)
    return clouds, np.array(ends), origin, dims


@pytest.mark.parametrize("seed,res,inflate,interval,thick", [(0, 0.4, 1, 1.0, 0.3), (1, 0.25, 2, 0.7, 0.2)])
def test_internal_space_and_outer_check_match_oracle(oracle_mod, seed, res, inflate, interval, thick):
    from fc_planner_b200.planner_grid import PlannerGrid
    clouds, ends, origin, dims = _skeleton_scene(seed, res)
    rng = np.random.default_rng(seed + 10)
    extra = np.concatenate(clouds)[::7] + rng.normal(0, 0.05, (len(np.concatenate(clouds)[::7]), 3)).astype(np.float32)   # "occ_model" denser cloud
    extra[:5] += 500.0                                                                                                      # outside the map: SetOcc skips them
    pg = PlannerGrid(origin, res, dims)
    og = oracle_mod.PlanOracle(origin, res, dims)
    for m in (pg, og):
        m.SetOcc(extra)
        m.InternalSpace(clouds, ends, inflate, interval, thick)
    for a, b, name in zip(pg.buffers(), og.buffers(), ("hc", "inflate", "internal")):
        assert np.array_equal(a, b), name
    assert pg.buffers()[2].sum() > 1000 and pg.buffers()[1].sum() > pg.buffers()[0].sum()
    # OuterCheck: rays from surface points outward (position, direction) clear internal marks
    allp = np.concatenate(clouds).astype(np.float64)
    c = allp.mean(0)
    dirs = allp - c
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    outers = np.concatenate([allp, dirs], axis=1)[::3]
    outers = np.concatenate([outers, np.concatenate([allp[:40], -dirs[:40]], axis=1)])      # some inward rays: they do clear marks
    before = pg.buffers()[2].sum()
    for m in (pg, og):
        m.OuterCheck(outers, 3.0)
    for a, b, name in zip(pg.buffers(), og.buffers(), ("hc", "inflate", "internal")):
        assert np.array_equal(a, b), name + " after OuterCheck"
    assert pg.buffers()[2].sum() < before
    assert pg.oob_writes() == og.oob_writes()
    assert pg.counters()[1] == 0 and og.counters()[1] == 0                                  # no march hit the step cap
    # the loops that read the buffers see the device-built planes
    P = np.concatenate([ends[:, :3], ends[:, 3:], allp[::200] + 3.0 * dirs[::200]])
    s1, d1 = pg.lineOfSight(P); s2, d2 = og.lineOfSight(P)
    assert np.array_equal(s1, s2) and np.array_equal(d1, d2) and 0 < s1.sum() < s1.size
    ok1, ok2 = pg.safetyBox(P.astype(np.float32), 1.0), og.safetyBox(P.astype(np.float32), 1.0)
    assert np.array_equal(ok1, ok2)


def test_device_built_grid_equals_uploaded_buffers(oracle_mod):
    """fcvm_grid_create from ready-made char buffers and the device-built grid give the same planes."""
    from fc_planner_b200.planner_grid import PlannerGrid
    clouds, ends, origin, dims = _skeleton_scene(3, 0.4)
    og = oracle_mod.PlanOracle(origin, 0.4, dims)
    og.InternalSpace(clouds, ends, 1, 1.0, 0.3)
    hc, inf, it = og.buffers()
    up = PlannerGrid(origin, 0.4, dims, hc, inf, it)
    for a, b in zip(up.buffers(), (hc, inf, it)):
        assert np.array_equal(a, b)
