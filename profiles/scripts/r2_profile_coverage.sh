# launch list of one CoverageEvaluation call on the MBS trajectory and a --set full capture of k_cov_scan, round 2
cat > /tmp/covprof.py <<'PY'
import os, numpy as np, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from conftest import load_golden, GOLDEN
from fc_planner_b200 import launch_params, launch_camera
from fc_planner_b200.coverage import CoverageEvaluator
_, zs = load_golden('mbs'); z = np.load(os.path.join(GOLDEN, 'cov_mbs.npz'))
ce = CoverageEvaluator(launch_params('mbs'), launch_camera('mbs')); ce.setCloud(zs['map_cloud'])
for it in range(3):
    r = ce.CoverageEvaluation(z['poses'])
print('scan ms', ce.last_scan_ms(), 'kernels ms', ce.last_kernel_ms())
PY
python /tmp/covprof.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${1:-r2z}_launches_cov.csv python /tmp/covprof.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cov_scan -s 2 -c 1 -o gpurun_out/${1:-r2z}_cov python /tmp/covprof.py > gpurun_out/ncu_cov.log 2>&1; tail -2 gpurun_out/ncu_cov.log
