for mode in 0 1 0 1; do
FCVM_PLAIN_ALLOC=$mode FCVM_PROFILE=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-prune --no-cov --no-plan --no-golden --no-stages 2> gpurun_out/full_$mode.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['pruning_full']; print('plain_alloc=$mode', f['seconds'], f['setInitViewpoints_s'], f['updateViewpoints_s'], f['first_call_seconds'])"
grep 'rank 0' gpurun_out/full_$mode.err | tail -1 | cut -c1-700
done
