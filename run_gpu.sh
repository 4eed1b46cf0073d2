set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1f.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','pruning','roofline','cpu_baseline')})
PY
tail -5 gpurun_out/bench_r1f.err
