python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -2 gpurun_out/bench_r1i.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1i.json').read().strip().splitlines()[-1])
print(f"{d['value']:.4g} rays/s {d['ms_per_step']:.1f} ms e2e {d['e2e']['value']:.4g} frac {d['roofline']['frac']:.3f}")
print(json.dumps(d.get('coverage_evaluation'), indent=1))
PY
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1i.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('pruning'), indent=1))
PY
