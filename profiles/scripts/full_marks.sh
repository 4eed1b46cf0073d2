# timeline marks of the full-scale updateViewpoints (configs[4]) over several repeats: where does the run-to-run spread come from?
for rep in 1 2 3; do
FCVM_PROFILE=3 python bench.py --steps 2 --warmup 3 --no-cpu --no-prune --no-cov --no-plan --no-golden --no-stages 2> gpurun_out/full_marks_$rep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=d['pruning_full']; print('rep $rep', f['seconds'], f['setInitViewpoints_s'], f['updateViewpoints_s'])"
grep -E 'updateViewpoints:|unc:|safety' gpurun_out/full_marks_$rep.err | tail -60 | cut -c1-150 > gpurun_out/full_marks_$rep.txt
done
