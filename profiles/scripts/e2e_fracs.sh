# e2e (fcvm_evaluate_csr with host buffers) against the sub-batch fractions of the CSR pipeline (FCVM_PIPE_FRACS, per cent)
for f in default 55,30,15 60,27,13 50,30,14,6 45,28,17,10 65,25,10 50,25,15,10 70,20,10; do
  if [ $f = default ]; then unset FCVM_PIPE_FRACS; else export FCVM_PIPE_FRACS=$f; fi
  python bench.py --steps 5 --warmup 3 --no-prune --no-cov --no-plan --no-cpu --no-full --no-golden --no-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fracs $f', 'step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
