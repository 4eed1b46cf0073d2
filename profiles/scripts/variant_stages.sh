# usage: bash profiles/scripts/variant_stages.sh default vA vB ...   (libraries under fc-planner_b200/_variants/): E1 step, e2e and E2-E4 kernel ms
for v in "$@"; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu --no-full --no-golden 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['ms_per_step'],2), 'kernel', round(d['roofline']['kernel_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:round(x['kernel_ms'],2) for k,x in d['other_evaluators_rank0'].items()})"
done
