"""fc-planner_b200/csrc/fcvm_ddmath.h on the CPU: the double-double asin / acos behind the device-side pitch / yaw sums.
The header is host + device code; here it is compiled with g++ and checked against libquadmath (accuracy) and against this
box's glibc (the certificate that lets the device skip libm: outside the flag margin the correctly rounded value IS glibc's)."""
import json
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _build():
    out = os.path.join(HERE, "_build", "ddmath_check")
    src = os.path.join(HERE, "ddmath_check.cpp")
    hdr = os.path.join(ROOT, "fc-planner_b200", "csrc", "fcvm_ddmath.h")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", src, "-o", out, "-lquadmath"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            pytest.skip("cannot build ddmath_check (libquadmath missing?): " + r.stdout[-400:])
    return out


@pytest.mark.parametrize("seed", [1, 2])
def test_dd_asin_acos_certificate(seed):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU without FMA")
    exe = _build()
    r = subprocess.run([exe, "1500000", str(seed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    z = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0, r.stderr[-2000:]
    # every deviation of glibc from correct rounding lies inside the flagged set ...
    assert z["unflagged_mismatch_asin"] == 0 and z["unflagged_mismatch_acos"] == 0
    # ... with room to spare: glibc deviates only within ~0.022 ulp of a midpoint, the flag margin is 0.03125 ulp
    assert z["worst_glibc_margin_asin"] < 0.025 and z["worst_glibc_margin_acos"] < 0.025
    # the double-double value itself is good to a small fraction of the margin
    assert z["max_err_ulp_asin"] < 1e-3 and z["max_err_ulp_acos"] < 1e-3
    # about 6.25 % of the arguments are flagged (2 x margin)
    assert 0.04 < z["flagged_asin"] / z["n"] < 0.09 and 0.04 < z["flagged_acos"] / z["n"] < 0.09


def test_division_from_reciprocal_is_exact(tmp_path):
    """div_exact (fcvm_device.cuh): reciprocal-table division with two fused multiply-adds == the hardware divider."""
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU without FMA")
    exe = str(tmp_path / "div_exact_check")
    r = subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", os.path.join(HERE, "div_exact_check.c"), "-o", exe, "-lm"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout
