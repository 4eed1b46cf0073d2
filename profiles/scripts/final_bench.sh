python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; tail -2 gpurun_out/bench_r1n.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1n.json').read().strip().splitlines()[-1])
print(f"{d['value']:.4g} rays/s {d['ms_per_step']:.1f} ms e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.1f} ms) rows {d['e2e_rows']['ms_per_step']:.1f} frac {d['roofline']['frac']:.3f} prune {d['pruning']['gpu_ms']:.1f} ms x{d['pruning']['speedup']:.0f} same {d['pruning']['selected_set_identical']} cov {d['coverage_evaluation']['gpu_ms']:.2f} ms x{d['coverage_evaluation']['speedup']:.0f} los {d['planner_grid']['los_kernel_ms']:.2f} ms cpu {d['cpu_baseline']['value']:.4g}")
PY
