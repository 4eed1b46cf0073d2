"""Hand-derived known answers for the oracle's primitives (SURVEY.md 8c micro-KATs): the values
below were worked out from the reference source by hand, not produced by the oracle.
References: plan_env/src/raycast.cpp:26-43 (signum/mod/intbound), 341-345 (setParams), 446-479
(nextId), 481-559 (biNextId); active_perception/src/perception_utils.cpp:115-124 (insideFOV);
viewpoint_manager.h:238-248 (PitchYaw)."""
import math

import numpy as np

from oracle import oracle


def test_intbound_known_answers():
    assert oracle.intbound(0.25, 1.0) == 0.75
    assert oracle.intbound(0.25, 2.0) == 0.375
    assert oracle.intbound(0.25, -1.0) == 0.25          # recursion: intbound(-0.25, 1) = (1 - 0.75) / 1
    assert oracle.intbound(-0.25, 1.0) == 0.25          # mod(-0.25, 1) = 0.75
    assert oracle.intbound(3.0, 2.0) == 0.5             # integral start, forward: one full cell
    assert oracle.intbound(3.0, -2.0) == 0.5            # integral start, backward: ALSO one full cell (quirk A.4 (3))
    assert oracle.intbound(0.25, 0.0) == math.inf       # idle axis
    assert oracle.intbound(0.25, -0.0) == -math.inf     # -0.0 < 0 is false: no recursion, 0.75 / -0.0 (unreachable: ds comes from an int)


def test_one_directional_march_axis_aligned():
    ids = oracle.ray_ids(1.0, [0, 0, 0], [0.5, 0.5, 0.5], [3.5, 0.5, 0.5])
    assert ids.tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]]      # last call returns false
    ids = oracle.ray_ids(1.0, [0, 0, 0], [3.5, 0.5, 0.5], [0.5, 0.5, 0.5])
    assert ids.tolist() == [[3, 0, 0], [2, 0, 0], [1, 0, 0], [0, 0, 0]]
    # start == end voxel: a single (false) call
    assert oracle.ray_ids(1.0, [0, 0, 0], [0.2, 0.2, 0.2], [0.8, 0.7, 0.6]).tolist() == [[0, 0, 0]]


def test_march_prefers_z_on_ties():
    # exact diagonal through cell centres: tMaxX == tMaxY == tMaxZ at every step; the ladder's final
    # `else` takes z, then (x vs y tie -> else branch) y, then x (raycast.cpp:455-476)
    ids = oracle.ray_ids(1.0, [0, 0, 0], [0.5, 0.5, 0.5], [2.5, 2.5, 2.5])
    assert ids.tolist() == [[0, 0, 0], [0, 0, 1], [0, 1, 1], [1, 1, 1], [1, 1, 2], [1, 2, 2], [2, 2, 2]]


def test_bidirectional_march_meets_in_the_middle():
    ids = oracle.biray_ids(1.0, [0, 0, 0], [0.5, 0.5, 0.5], [4.5, 0.5, 0.5])
    assert ids.tolist() == [[0, 0, 0, 4, 0, 0], [1, 0, 0, 3, 0, 0], [2, 0, 0, 2, 0, 0]]
    # uneven halves: the p side (2 voxels from the middle) keeps going after the n side (1 voxel) has arrived
    ids = oracle.biray_ids(1.0, [0, 0, 0], [0.5, 0.5, 0.5], [3.9, 0.5, 0.5])   # mid = 2.2 -> voxel 2
    assert ids.tolist() == [[0, 0, 0, 3, 0, 0], [1, 0, 0, 2, 0, 0], [2, 0, 0, 2, 0, 0]]


def test_map_index_truncates_toward_zero():
    # origin (1,1,1), res 1: offset_ = 0.5 - 1 = -0.5; world voxel v -> (int)(v - 0.5):
    # v = 2 -> 1, v = 1 -> 0, v = 0 -> (int)(-0.5) = 0 (NOT -1: quirk A.4 (1)), v = -1 -> -1
    ids = oracle.ray_ids(1.0, [1, 1, 1], [2.5, 1.5, 1.5], [-1.5, 1.5, 1.5])
    assert ids[:, 0].tolist() == [1, 0, 0, -1, -2]
    assert (ids[:, 1:] == 0).all()


def test_inside_fov_known_answers(small_plant):
    prm = small_plant["params"]            # max_dist 11, fov 55 x 75 deg with pi = 3.1415926
    pose = [0, 0, 0, 0, 0]                 # looking along +x
    top = prm.top_angle
    left = prm.left_angle
    pts = [
        [5, 0, 0],                                          # on the axis
        [11, 0, 0],                                         # exactly at max_dist: `norm > max_dist` is false
        [np.nextafter(11.0, 12.0), 0, 0],                   # one ulp beyond
        [-1, 0, 0],                                         # behind
        [5 * math.cos(top - 1e-6), 0, 5 * math.sin(top - 1e-6)],     # just inside the top plane
        [5 * math.cos(top + 1e-6), 0, 5 * math.sin(top + 1e-6)],     # just outside
        [5 * math.cos(left - 1e-6), 5 * math.sin(left - 1e-6), 0],   # just inside the left plane
        [5 * math.cos(left + 1e-6), -5 * math.sin(left + 1e-6), 0],  # just outside the right plane
        [0, 0, 0],                                          # the camera centre itself: dir = 0, all dots = 0 >= 0
    ]
    assert oracle.inside_fov(prm, pose, pts).tolist() == [True, True, False, False, True, False, True, False, True]
    # yaw = pi/2 looks along +y
    assert oracle.inside_fov(prm, [0, 0, 0, 0, math.pi / 2], [[0, 5, 0], [5, 0, 0]]).tolist() == [True, False]
    # pitch > 0 looks up (R_pitch maps body x to (cos p, 0, sin p))
    assert oracle.inside_fov(prm, [0, 0, 0, 1.0, 0], [[5 * math.cos(1.0), 0, 5 * math.sin(1.0)], [5, 0, 0]]).tolist() == [True, False]


def test_pitch_yaw_known_answers():
    assert oracle.pitch_yaw([0, 0, 1]).tolist() == [math.asin(1 / (1 + 1e-3)), 0.0]
    assert oracle.pitch_yaw([0, 2, 0]).tolist() == [0.0, math.atan2(2, 0)]
    assert oracle.pitch_yaw([-1, 0, 0]).tolist() == [0.0, math.pi]
