timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -2 gpurun_out/bench_r1j.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1j.json').read().strip().splitlines()[-1])
print(f"{d['value']:.4g} rays/s {d['ms_per_step']:.1f} ms e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.1f} ms) frac {d['roofline']['frac']:.3f} prune {d['pruning']['gpu_ms']:.1f} ms x{d['pruning']['speedup']:.0f} same {d['pruning']['selected_set_identical']} cov {d['coverage_evaluation']['gpu_ms']:.2f} ms")
PY
