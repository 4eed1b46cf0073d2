"""The C++ hooks header for the widened rows (fc-planner_b200/shim/hierarchical_coverage_planner/fcvm_planner_hooks.h) compiles
against stand-in ROS / PCL / Eigen headers, links to libfcvm.so and runs the planner-side loops with the reference's types.
Without a GPU it must refuse by logging (no exception, empty outputs); with a GPU its outputs must equal the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

from test_oracle_plan import _grid, _segments

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
BIN = os.path.join(HERE, "_build", "hooks_harness")


def _build():
    from fc_planner_b200 import _lib
    _lib.load()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-O2", "-std=c++14", "-Wall", "-I", os.path.join(HERE, "shim_stubs"), "-I", os.path.join(ROOT, "fc-planner_b200", "shim"),
           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "hooks_harness.cpp"), "-o", BIN, "-L", libdir, "-lfcvm", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return BIN


def _inputs(small_plant, tmp):
    from fc_planner_b200 import launch_camera, scenes
    prm, cam = small_plant["params"], launch_camera("mbs")
    cloud = small_plant["points"]
    poses = scenes.poses_from_viewpoints(small_plant["viewpoints"][:120])
    origin, res, dims, hc, inflate, _ = _grid(5, (30, 28, 26), 0.4, fill=0.04)
    rng = np.random.default_rng(9)
    a, b = _segments(rng, origin, res, dims, 60)
    rosa = b[:25].astype(np.float32)
    with open(os.path.join(tmp, "in.bin"), "wb") as f:
        f.write(struct.pack("<4d", prm.top_angle, prm.left_angle, prm.right_angle, prm.max_dist))
        f.write(struct.pack("<6d", cam.fx, cam.fy, cam.cx, cam.cy, cam.fov_h, cam.fov_w))
        f.write(struct.pack("<4q", len(cloud), len(poses), len(rosa), len(a)))
        f.write(np.ascontiguousarray(cloud, np.float32).tobytes()); f.write(np.ascontiguousarray(poses, np.float64).tobytes())
        f.write(np.ascontiguousarray(origin, np.float64).tobytes()); f.write(struct.pack("<d", res)); f.write(np.ascontiguousarray(dims, np.int32).tobytes())
        f.write(hc.astype(np.int8).tobytes()); f.write(inflate.astype(np.int8).tobytes())
        f.write(rosa.tobytes()); f.write(np.ascontiguousarray(a, np.float64).tobytes())
    return prm, cam, cloud, poses, (origin, res, dims, hc, inflate), rosa, a


def _run(exe, tmp):
    return subprocess.run([exe, os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)


def test_hooks_compile_link_and_refuse_without_device(small_plant, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; see the gpu test")
    exe = _build()
    _, _, _, poses, _, _, a = _inputs(small_plant, str(tmp_path))
    r = _run(exe, str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    assert struct.unpack_from("<d", raw, 0)[0] == 0.0
    assert not any(struct.unpack_from(f"<{len(poses)}q", raw, 8))          # every visible group empty


@pytest.mark.gpu
def test_hooks_match_oracle(small_plant, oracle_mod, tmp_path):
    exe = _build()
    prm, cam, cloud, poses, grid, rosa, a = _inputs(small_plant, str(tmp_path))
    r = _run(exe, str(tmp_path))
    assert r.returncode == 0 and r.stderr.strip() == "", r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    m, n = len(poses), len(a)
    rate = struct.unpack_from("<d", raw, 0)[0]
    sizes = np.frombuffer(raw, "<i8", m, 8)
    off = 8 + 8 * m
    pts = np.frombuffer(raw, "<f4", 3 * int(sizes.sum()), off).reshape(-1, 3)
    off += 12 * int(sizes.sum())
    path_cov = struct.unpack_from("<d", raw, off)[0]; off += 8
    det = np.frombuffer(raw, np.uint8, n, off); off += n
    box = np.frombuffer(raw, np.uint8, n, off); off += n
    safe = np.frombuffer(raw, np.uint8, n * n, off).reshape(n, n); off += n * n
    dist = np.frombuffer(raw, "<f8", n * n, off).reshape(n, n)
    oce = oracle_mod.CoverageOracle(prm, cam)
    oce.setCloud(cloud)
    o_rate, o_offs, o_ids, _ = oce.CoverageEvaluation(poses)
    assert rate == o_rate and np.array_equal(sizes, np.diff(o_offs)) and np.array_equal(pts, cloud[o_ids]) and len(o_ids) > 100
    _, _, o_cov, _ = oce.pinhole(poses)
    assert path_cov == float(o_cov.sum()) / len(o_cov)
    origin, res, dims, hc, inflate = grid
    opg = oracle_mod.PlanOracle(origin, res, dims, hc, inflate, None)
    o_det, _, _ = opg.lightStateDet(a, rosa)
    assert np.array_equal(det, o_det)
    assert np.array_equal(box, opg.safetyBox(a.astype(np.float32), 1.0))
    o_safe, o_dist = opg.lineOfSight(a)
    assert np.array_equal(safe, o_safe) and np.array_equal(dist, o_dist)
