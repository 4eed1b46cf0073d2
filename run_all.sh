timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for st in 1 3; do python bench.py --stage $st --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('E$st', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['roofline']['frac'])"; done
