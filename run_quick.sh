for n in 4 6 8 12 16; do FCVM_PIPE_SUBS=$n python bench.py --steps 3 --warmup 3 --no-prune --no-cov --no-plan --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('subs $n', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'])"; done
