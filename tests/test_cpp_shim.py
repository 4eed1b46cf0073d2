"""The C++ drop-in header (fc-planner_b200/shim/viewpoint_manager/viewpoint_manager.h) compiles
against stand-in ROS / PCL / Eigen headers, links to libfcvm.so, and runs the reference's call
sequence.  Without a GPU the library must refuse (log + return, like the reference's error
convention) and the outputs stay empty; with a GPU the result must equal the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
BIN = os.path.join(HERE, "_build", "shim_harness")


def _build():
    from fc_planner_b200 import _lib
    _lib.load()
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.dirname(_lib.lib_path())
    cmd = ["g++", "-O2", "-std=c++14", "-Wall", "-I", os.path.join(HERE, "shim_stubs"), "-I", os.path.join(ROOT, "fc-planner_b200", "shim"),
           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "shim_harness.cpp"), "-o", BIN, "-L", libdir, "-lfcvm",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return BIN


def _write_inputs(tmp, prm, pts, nrm, vps, gpus=1, drop=()):
    keys = {"viewpoint_manager/visible_range": prm.visible_range, "viewpoint_manager/viewpoints_distance": prm.viewpoints_distance,
            "viewpoint_manager/fov_h": prm.fov_h, "viewpoint_manager/fov_w": prm.fov_w, "viewpoint_manager/pitch_upper": prm.pitch_upper,
            "viewpoint_manager/pitch_lower": prm.pitch_lower, "viewpoint_manager/zGround": prm.z_ground, "viewpoint_manager/GroundPos": prm.ground_pos,
            "viewpoint_manager/safeHeight": prm.safe_height, "viewpoint_manager/safe_radius": prm.safe_radius,
            "viewpoint_manager/max_iter_num": prm.max_iter_num, "viewpoint_manager/pose_update": prm.pose_update,
            "perception_utils/top_angle": prm.top_angle, "perception_utils/left_angle": prm.left_angle, "perception_utils/right_angle": prm.right_angle,
            "perception_utils/max_dist": prm.max_dist, "perception_utils/vis_dist": prm.vis_dist, "hcmap/resolution": prm.resolution,
            "fcvm/seed": prm.seed, "fcvm/gpus": gpus}
    for k in drop:
        keys.pop(k)
    with open(os.path.join(tmp, "params.txt"), "w") as f:
        for k, v in keys.items():
            f.write(f"{k} {float(v)!r}\n")
        f.write("viewpoint_manager/attitude_type %s\n" % {1: "yaw", 2: "all"}.get(prm.attitude, "null"))
    with open(os.path.join(tmp, "input.bin"), "wb") as f:
        f.write(struct.pack("<3q", len(pts), len(pts), len(vps)))
        f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        f.write(np.ascontiguousarray(nrm, np.float64).tobytes())
        f.write(np.ascontiguousarray(vps, np.float32).tobytes())


def _read_outputs(path):
    raw = open(path, "rb").read()
    nf = struct.unpack_from("<q", raw, 0)[0]
    rec = np.dtype([("id", "<i4"), ("pose", "<f8", 5), ("vox", "<i4")])
    fin = np.frombuffer(raw, rec, nf, 8)
    off = 8 + nf * rec.itemsize
    nu = struct.unpack_from("<q", raw, off)[0]
    unc = np.frombuffer(raw, "<f4", 3 * nu, off + 8).reshape(-1, 3)
    return fin, unc


def test_shim_compiles_links_and_refuses_without_device(small_plant, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present; see the gpu test")
    exe = _build()
    _write_inputs(str(tmp_path), small_plant["params"], small_plant["points"], small_plant["normals"], small_plant["viewpoints"][:50])
    r = subprocess.run([exe, str(tmp_path / "params.txt"), str(tmp_path / "input.bin"), str(tmp_path / "out.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr          # errors are logged, never thrown (the reference's convention)
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr
    fin, unc = _read_outputs(tmp_path / "out.bin")
    assert len(fin) == 0 and len(unc) == 0


@pytest.mark.gpu
def test_shim_sequence_matches_oracle(small_plant, oracle_mod, tmp_path):
    exe = _build()
    prm = small_plant["params"]
    vps = small_plant["viewpoints"][:200]
    _write_inputs(str(tmp_path), prm, small_plant["points"], small_plant["normals"], vps)
    r = subprocess.run([exe, str(tmp_path / "params.txt"), str(tmp_path / "input.bin"), str(tmp_path / "out.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0 and r.stderr.strip() == "", r.stderr
    fin, unc = _read_outputs(tmp_path / "out.bin")
    orc = oracle_mod.Oracle(prm)
    orc.setMapPointCloud(small_plant["points"]); orc.setModel(small_plant["points"])
    orc.setNormals(small_plant["points"].astype(np.float64), small_plant["normals"])
    orc.setInitViewpoints(vps)
    assert orc.updateViewpoints() == 0
    oi, op, ox = orc.getUpdatedViewpoints()
    assert len(oi) > 5
    assert np.array_equal(fin["id"], oi) and np.array_equal(fin["pose"], op) and np.array_equal(fin["vox"], ox)
    assert np.array_equal(unc, orc.getUncoveredModelCloud())


@pytest.mark.gpu
def test_shim_on_two_gpus_in_one_process(small_plant, oracle_mod, tmp_path):
    """`fcvm/gpus 2`: the drop-in class holds one library handle per GPU, joined by the library's own NCCL communicator and
    driven by one host thread each; the planner's call sequence is unchanged and the result is the oracle's, bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = _build()
    prm = small_plant["params"]
    vps = small_plant["viewpoints"][:301]
    _write_inputs(str(tmp_path), prm, small_plant["points"], small_plant["normals"], vps, gpus=2)
    r = subprocess.run([exe, str(tmp_path / "params.txt"), str(tmp_path / "input.bin"), str(tmp_path / "out.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and "ERROR" not in r.stderr, r.stderr
    fin, unc = _read_outputs(tmp_path / "out.bin")
    orc = oracle_mod.Oracle(prm)
    orc.setMapPointCloud(small_plant["points"]); orc.setModel(small_plant["points"])
    orc.setNormals(small_plant["points"].astype(np.float64), small_plant["normals"])
    orc.setInitViewpoints(vps)
    assert orc.updateViewpoints() == 0
    oi, op, ox = orc.getUpdatedViewpoints()
    assert len(oi) > 5
    assert np.array_equal(fin["id"], oi) and np.array_equal(fin["pose"], op) and np.array_equal(fin["vox"], ox)
    assert np.array_equal(unc, orc.getUncoveredModelCloud())


@pytest.mark.gpu
def test_shim_defaults_are_the_references(small_plant, oracle_mod, tmp_path):
    """viewpoint_manager/max_iter_num and viewpoint_manager/pose_update left unset: the reference defaults to 0 and TRUE
    (viewpoint_manager.cpp:46-47), so the E1 / E2 pose refinement runs and no uncovered-area round does."""
    exe = _build()
    prm = small_plant["params"]
    vps = small_plant["viewpoints"][:150]
    _write_inputs(str(tmp_path), prm, small_plant["points"], small_plant["normals"], vps, drop=("viewpoint_manager/max_iter_num", "viewpoint_manager/pose_update"))
    r = subprocess.run([exe, str(tmp_path / "params.txt"), str(tmp_path / "input.bin"), str(tmp_path / "out.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0 and r.stderr.strip() == "", r.stderr
    fin, unc = _read_outputs(tmp_path / "out.bin")
    import copy
    p2 = copy.copy(prm)
    p2.max_iter_num, p2.pose_update = 0, 1
    orc = oracle_mod.Oracle(p2)
    orc.setMapPointCloud(small_plant["points"]); orc.setModel(small_plant["points"])
    orc.setNormals(small_plant["points"].astype(np.float64), small_plant["normals"])
    orc.setInitViewpoints(vps)
    assert orc.updateViewpoints() == 0
    oi, op, ox = orc.getUpdatedViewpoints()
    assert np.array_equal(fin["id"], oi) and np.array_equal(fin["pose"], op) and np.array_equal(fin["vox"], ox)
