import sys, os
sys.path.insert(0, os.getcwd())
from fc_planner_b200 import _lib
for k in [k for k in _lib.SYMBOLS if k.startswith('fcvm_cov_') or k == 'fcvm_safety_check']:
    del _lib.SYMBOLS[k]
sys.argv = ['bench.py', '--steps', '3', '--warmup', '3', '--no-prune', '--no-cov', '--no-cpu']
import bench
bench.main()
