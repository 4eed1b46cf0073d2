"""Parameters of the viewpoint-manager path, as the reference reads them from the ROS
parameter server (viewpoint_manager.cpp:35-47, perception_utils.cpp:32-37, sdf_map.cpp:128-136),
and the values its launch files give them.

`Params` is the ctypes image of `fcvm_params` in include/fcvm.h (the oracle uses the same layout).
`launch_params(scene)` reproduces the doubles the launch XML produces, including the
`$(eval arg('fov_h')*3.1415926/360.0)` arithmetic with its truncated pi
(viewpoint_manager/launch/mbs.launch:40-42) — Python float semantics, left to right.
"""
import ctypes

ATTITUDE_OTHER, ATTITUDE_YAW, ATTITUDE_ALL = 0, 1, 2


class Params(ctypes.Structure):
    _fields_ = [
        ("visible_range", ctypes.c_double),
        ("viewpoints_distance", ctypes.c_double),
        ("fov_h", ctypes.c_double),
        ("fov_w", ctypes.c_double),
        ("pitch_upper", ctypes.c_double),
        ("pitch_lower", ctypes.c_double),
        ("ground_pos", ctypes.c_double),
        ("safe_height", ctypes.c_double),
        ("safe_radius", ctypes.c_double),
        ("top_angle", ctypes.c_double),
        ("left_angle", ctypes.c_double),
        ("right_angle", ctypes.c_double),
        ("max_dist", ctypes.c_double),
        ("vis_dist", ctypes.c_double),
        ("resolution", ctypes.c_double),
        ("z_ground", ctypes.c_int32),
        ("attitude", ctypes.c_int32),
        ("max_iter_num", ctypes.c_int32),
        ("pose_update", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("device", ctypes.c_int32),
        ("host_threads", ctypes.c_int32),
    ]

    def copy(self, **changes):
        q = Params()
        ctypes.memmove(ctypes.byref(q), ctypes.byref(self), ctypes.sizeof(Params))
        for k, v in changes.items():
            setattr(q, k, v)
        return q

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class Camera(ctypes.Structure):
    """ctypes image of `fcvm_camera`: hcplanner/{fx, fy, cx, cy} (hcplanner.cpp:56-59) and the degrees
    viewpoint_manager/{fov_h, fov_w} that CoverageEvaluation / PinHoleCamera combine with M_PI (:1040-1041)."""
    _fields_ = [("fx", ctypes.c_double), ("fy", ctypes.c_double), ("cx", ctypes.c_double), ("cy", ctypes.c_double),
                ("fov_h", ctypes.c_double), ("fov_w", ctypes.c_double)]


def launch_camera(scene="mbs"):
    """hierarchical_coverage_planner/launch/{mbs,pipe,christ}.launch:5-8 (identical in all three)."""
    s = _SCENES[scene]
    return Camera(fx=417.0467, fy=461.0951, cx=320.0, cy=240.0, fov_h=s["fov_h"], fov_w=s["fov_w"])


# scene -> (fov_h, fov_w, max_dist, resolution, dist_vp, pitch_upper, pitch_lower, zGround, GroundPos,
#           safeHeight, safe_radius, attitude, max_iter_num, pose_update, model leaf)
# viewpoint_manager/launch/{mbs,pipe,christ}.launch; hierarchical_coverage_planner/launch/* carry the
# same values for these keys.
_SCENES = {
    "mbs":    dict(fov_h=55.0, fov_w=75.0, max_dist=11.0, resolution=0.4, dist_vp=8.0, pitch_upper=70.0, pitch_lower=-85.0,
                   z_ground=1, ground_pos=-10.5, safe_height=1.5, safe_radius=3.0, attitude=ATTITUDE_ALL, max_iter_num=3,
                   pose_update=1, leaf=2.0),
    "pipe":   dict(fov_h=55.0, fov_w=75.0, max_dist=11.0, resolution=0.4, dist_vp=8.0, pitch_upper=70.0, pitch_lower=-85.0,
                   z_ground=0, ground_pos=-7.4, safe_height=1.5, safe_radius=3.0, attitude=ATTITUDE_ALL, max_iter_num=5,
                   pose_update=1, leaf=2.0),
    "christ": dict(fov_h=55.0, fov_w=75.0, max_dist=8.0, resolution=0.4, dist_vp=6.0, pitch_upper=70.0, pitch_lower=-85.0,
                   z_ground=0, ground_pos=-7.4, safe_height=1.5, safe_radius=3.0, attitude=ATTITUDE_ALL, max_iter_num=2,
                   pose_update=1, leaf=0.8),
}
# the synthetic scaled scene of BASELINE.json configs[4]: MBS parameters, no ground, 0.1 m grid
_SCENES["synthetic"] = dict(_SCENES["mbs"], z_ground=0, resolution=0.1, leaf=None)


def launch_params(scene="mbs", seed=0, device=0, **overrides):
    s = dict(_SCENES[scene])
    s.update({k: v for k, v in overrides.items() if k in s})
    p = Params()
    p.visible_range = s["max_dist"]
    p.viewpoints_distance = s["dist_vp"]
    p.fov_h, p.fov_w = s["fov_h"], s["fov_w"]
    p.pitch_upper, p.pitch_lower = s["pitch_upper"], s["pitch_lower"]
    p.ground_pos, p.safe_height, p.safe_radius = s["ground_pos"], s["safe_height"], s["safe_radius"]
    p.top_angle = s["fov_h"] * 3.1415926 / 360.0
    p.left_angle = s["fov_w"] * 3.1415926 / 360.0
    p.right_angle = s["fov_w"] * 3.1415926 / 360.0
    p.max_dist = s["max_dist"]
    p.vis_dist = 3.0
    p.resolution = s["resolution"]
    p.z_ground = s["z_ground"]
    p.attitude = s["attitude"]
    p.max_iter_num = s["max_iter_num"]
    p.pose_update = s["pose_update"]
    p.seed = seed
    p.device = device
    p.host_threads = 0
    for k, v in overrides.items():
        if k not in s and hasattr(p, k):
            setattr(p, k, v)
    return p


def model_leaf(scene):
    return _SCENES[scene]["leaf"]
