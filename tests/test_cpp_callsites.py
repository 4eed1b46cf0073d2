"""'hcplanner.cpp, skeleton_viewpoint.cpp and viewpoint_manager_node.cpp compile unchanged' (INTEGRATION.md), shown rather
than asserted: the three call blocks of predrecon::ViewpointManager are cut out of the reference sources where they lie
(/root/reference, present in the build container only) and compiled byte for byte against the drop-in header and
libfcvm.so.  Nothing of the reference is stored in the repository; without /root/reference the test is skipped."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
REF = "/root/reference/FC-Planner/src"
BLOCKS = {   # file, first line, last line, first and last statement expected inside
    "hcplanner": ("hierarchical_coverage_planner/src/hcplanner.cpp", 708, 716, "vector<SingleViewpoint> updated_vps;", "getUncoveredModelCloud(uncovered_area);"),
    "skeleton": ("viewpoint_manager/src/skeleton_viewpoint.cpp", 488, 496, "vector<SingleViewpoint> updated_vps;", "getUncoveredModelCloud(uncovered_area);"),
    "node": ("viewpoint_manager/src/viewpoint_manager_node.cpp", 129, 136, "vector<SingleViewpoint> updated_vps;", "getUpdatedViewpoints(updated_vps);"),
}


def test_reference_call_sites_compile_unchanged_against_the_shim():
    if not os.path.isdir(REF):
        pytest.skip("/root/reference is not present on this box")
    from fc_planner_b200 import _lib
    _lib.load()
    build = os.path.join(HERE, "_build")
    os.makedirs(build, exist_ok=True)
    for name, (rel, lo, hi, first, last) in BLOCKS.items():
        lines = open(os.path.join(REF, rel)).read().splitlines()[lo - 1:hi]
        assert first in lines[0] and last in lines[-1], (name, lines[0], lines[-1])
        calls = [ln.split("->")[1].split("(")[0] for ln in lines if "viewpoint_manager_->" in ln]
        assert calls[:6] == ["reset", "setMapPointCloud", "setModel", "setNormals", "setInitViewpoints", "updateViewpoints"], calls
        with open(os.path.join(build, f"callsite_{name}.inc"), "w") as f:
            f.write("\n".join(lines) + "\n")
    exe = os.path.join(build, "ref_callsites")
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-O1", "-std=c++14", "-Wall", "-I", build, "-I", os.path.join(HERE, "shim_stubs"), "-I", os.path.join(ROOT, "fc-planner_b200", "shim"),
           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "ref_callsite_tu.cpp"), "-o", exe, "-L", libdir, "-lfcvm", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "callsites ran: 0 viewpoints" in r.stdout          # empty inputs: logged and refused, as the reference does
    assert "Input" in r.stderr or "empty" in r.stderr
