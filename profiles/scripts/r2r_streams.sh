timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_gputest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
for f in "40,25,15,12,8" "35,25,20,12,8" "45,30,15,10" "30,25,20,15,10"; do
  echo "fracs $f" >> gpurun_out/r2s_fracs.txt
  FCVM_PIPE_FRACS=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-prune --no-cov --no-plan --no-cpu --no-full --no-golden --no-stages 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))" >> gpurun_out/r2s_fracs.txt 2>&1
done
tail -3 gpurun_out/r2s_gputest.log; cat gpurun_out/r2s_fracs.txt; cut -c1-400 gpurun_out/r2s_bench.json
