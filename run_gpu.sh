set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1e.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','pruning')})
PY
tail -5 gpurun_out/bench_r1e.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-prune > gpurun_out/plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o gpurun_out/prof_r1e python bench.py --steps 2 --warmup 1 --no-cpu --no-prune > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log | cut -c1-200
