for v in exp1 exp2; do
  export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so
  python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'])
except Exception as e: print('$v failed', e)"
done
