for v in w3 w6 w8; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  FCVM_PIPE_SUBS=2 python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], 'csr', d['e2e']['ms_per_step'], 'rows', d['e2e_rows']['ms_per_step'])"; done
