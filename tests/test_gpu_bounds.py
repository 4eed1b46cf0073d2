"""Memory safety of k_eval without compute-sanitizer (the tool is closed on this GPU pool): the
FCVM_DEBUG_BOUNDS build checks every data-dependent index (queue slots, tile offsets, bitset words,
brick indices) and counts violations; a run over all stages, the full call sequence, lattice-aligned
viewpoints (runaway marcher) and viewpoints outside the map must count none."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_index_checking_build_sees_no_violation():
    dbg = os.path.join(ROOT, "fc-planner_b200", "_build", "libfcvm_dbg.so")
    if not os.path.exists(dbg):
        pytest.skip("libfcvm_dbg.so not built (python __graft_entry__.py)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sanitize_small.py")], env=dict(os.environ, FCVM_LIB=dbg),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "bounds violations 0\n" in r.stdout, r.stdout[-3000:]
    assert "sanitize_small ok" in r.stdout
