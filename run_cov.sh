timeout 900 python -m pytest tests/test_gpu_cov.py -x -q 2>&1 | tail -5
cat > /tmp/covtime.py <<'PY'
import os, time, numpy as np, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from conftest import load_golden, GOLDEN
from fc_planner_b200 import launch_params, launch_camera
from fc_planner_b200.coverage import CoverageEvaluator
for scene in ('mbs', 'pipe', 'christ'):
    _, zs = load_golden(scene); z = np.load(os.path.join(GOLDEN, f'cov_{scene}.npz'))
    ce = CoverageEvaluator(launch_params(scene), launch_camera(scene)); ce.setCloud(zs['map_cloud'])
    for it in range(3):
        t = time.time(); r = ce.CoverageEvaluation(z['poses']); w = time.time() - t
    print(os.environ.get('FCVM_LIB', 'default')[-14:], scene, 'evaluate wall ms', round(w * 1e3, 2), 'kernels ms', round(ce.last_kernel_ms(), 3), 'scan ms', round(ce.last_scan_ms(), 3), 'rate', r[0])
PY
timeout 300 python /tmp/covtime.py 2>&1 | tail -4
for c in 2 3; do FCVM_LIB=$PWD/fc-planner_b200/_variants/cov_ctas$c.so timeout 300 python /tmp/covtime.py 2>&1 | tail -3; done
