timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bounds.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for v in default fcmp; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  for st in 1 4; do python bench.py --stage $st --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v E$st', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; done
done
