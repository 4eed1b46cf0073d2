# usage: bash profiles/scripts/variant_bench.sh default qA qB ...   (libraries under fc-planner_b200/_variants/)
for v in "$@"; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])"
done
