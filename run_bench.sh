python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-cpu > gpurun_out/bench_plan.json 2> gpurun_out/bench_plan.err; tail -3 gpurun_out/bench_plan.err
python bench.py --steps 1 --warmup 3 --no-prune --no-cov --cpu-seconds 2 > gpurun_out/bench_plan2.json 2> gpurun_out/bench_plan2.err; tail -3 gpurun_out/bench_plan2.err
python - <<'PY'
import json
for f in ('bench_plan','bench_plan2'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(json.dumps(d.get('planner_grid'), indent=1))
PY
