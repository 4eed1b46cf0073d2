"""The C-ABI library loads without a GPU, exports every symbol include/fcvm.h declares, its
host-only helpers agree with the oracle bit for bit, and every evaluating entry point fails loudly
(FCVM_ENODEVICE) instead of falling back to a CPU implementation.  No device compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fcvm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fcvm_[a-z0-9_]+)\s*\(", src)) - {"fcvm_comm_fn"})


def test_library_exports_every_declared_symbol():
    from fc_planner_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libfcvm.so does not export {n}"
        assert n in _lib.SYMBOLS, f"ctypes mirror lacks {n}"
    assert set(_lib.SYMBOLS) <= set(names)
    assert lib.fcvm_abi_version() == 2


def test_params_struct_layout_matches_header():
    from fc_planner_b200 import Params
    src = open(os.path.join(ROOT, "include", "fcvm.h")).read()
    body = re.search(r"typedef struct fcvm_params \{(.*?)\} fcvm_params;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        typ, names = decl.split(None, 1)
        fields += [(n.strip(), typ) for n in names.split(",")]
    ctype = {"double": ctypes.c_double, "int32_t": ctypes.c_int32, "uint64_t": ctypes.c_uint64}
    assert [(n, ctype[t]) for n, t in fields] == list(Params._fields_)


def test_host_fov_normals_match_oracle(oracle_mod, small_plant):
    from fc_planner_b200.manager import host_fov_normals
    prm = small_plant["params"]
    rng = np.random.default_rng(0)
    for _ in range(300):
        p, y = rng.uniform(-1.6, 1.6), rng.uniform(-4, 4)
        assert np.array_equal(host_fov_normals(prm, p, y), oracle_mod.fov_normals(prm, p, y))


@pytest.mark.parametrize("scene", ["christ", "plant"])
def test_host_grid_matches_oracle(oracle_mod, small_plant, scene):
    """initHCMap geometry and voxelisation (sdf_map.cpp:122-188): origin, voxel counts and every
    occupancy bit, exported from the bricked device layout back to the reference's x-major order."""
    from fc_planner_b200.manager import host_build_grid
    if scene == "plant":
        prm, cloud = small_plant["params"], small_plant["points"]
    else:
        from conftest import load_golden
        prm, z = load_golden("christ")
        cloud = z["map_cloud"]
    orc = oracle_mod.Oracle(prm)
    orc.setMapPointCloud(cloud)
    o_origin, o_dims, o_bits = orc.grid()
    origin, dims, bits = host_build_grid(prm, cloud)
    assert np.array_equal(origin, o_origin) and np.array_equal(dims, o_dims)
    assert np.array_equal(bits, o_bits)
    assert np.unpackbits(bits).sum() > 1000


def test_shard_range_partitions_everything():
    from fc_planner_b200.manager import shard_range
    for n in (0, 1, 7, 8, 100, 200_000):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover += list(range(lo, hi)) if n <= 100 else []
                if r:
                    assert lo == prev_hi
                prev_hi = hi
            assert prev_hi == n
            if n <= 100:
                assert cover == list(range(n))


def test_no_cpu_fallback_without_device(small_plant):
    """On a box without a GPU every device-touching call must fail with FCVM_ENODEVICE (-3)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fc_planner_b200 import _lib
    from fc_planner_b200.manager import ViewpointManager
    vm = ViewpointManager(small_plant["params"])
    with pytest.raises(_lib.FcvmError) as e:
        vm.setMapPointCloud(small_plant["points"])
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.FcvmError) as e:
        vm.evaluate(1, np.zeros((1, 5)))
    assert e.value.code in (-2, -3)


def test_widened_entry_points_have_no_cpu_fallback_either(small_plant):
    """fcvm_cov_* and fcvm_grid_* (SURVEY.md 8f): creating a coverage handle is host-only, every evaluating call and the
    grid upload need the device and say so."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from fc_planner_b200 import _lib, launch_camera
    from fc_planner_b200.coverage import CoverageEvaluator
    from fc_planner_b200.planner_grid import PlannerGrid
    ce = CoverageEvaluator(small_plant["params"], launch_camera("mbs"))
    with pytest.raises(_lib.FcvmError) as e:
        ce.setCloud(small_plant["points"])
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.FcvmError) as e:
        ce.PinHoleCamera(np.zeros((1, 5)))
    assert e.value.code == -2                      # no cloud
    with pytest.raises(_lib.FcvmError) as e:
        PlannerGrid([0, 0, 0], 0.5, [8, 8, 8])
    assert "no CPU fallback" in str(e.value)


def test_camera_struct_layout_matches_header():
    from fc_planner_b200 import Camera
    src = open(os.path.join(ROOT, "include", "fcvm.h")).read()
    body = re.search(r"typedef struct fcvm_camera \{(.*?)\} fcvm_camera;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip() for decl in body.split(";") if decl.strip() for n in decl.strip().split(None, 1)[1].split(",")]
    assert names == [f for f, _ in Camera._fields_] and all(t is ctypes.c_double for _, t in Camera._fields_)


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under fc-planner_b200/ may import, link or open it."""
    pkg = os.path.join(ROOT, "fc-planner_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "fcvm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
