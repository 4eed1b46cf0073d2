for v in default V1 V2 V3 V4; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"
done
