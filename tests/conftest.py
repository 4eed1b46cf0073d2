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
