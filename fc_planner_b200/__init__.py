"""Import name of the `fc-planner_b200/` package (a hyphen cannot be imported).

The sources live in `fc-planner_b200/`; this stub only extends the package path there.
"""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fc-planner_b200"))

from .params import Camera, Params, launch_camera, launch_params  # noqa: E402,F401
