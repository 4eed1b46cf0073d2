timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; tail -3 gpurun_out/b.err; python -c "
import json
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'], d['e2e_rows'])"
FCVM_PROFILE=1 python profiles/fullscale_update.py 25000 2>&1 | tail -3
