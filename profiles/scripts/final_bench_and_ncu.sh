python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -2 gpurun_out/bench_r1l.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1l_ref.json 2> gpurun_out/bench_r1l_ref.err; tail -2 gpurun_out/bench_r1l_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1l.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 3 -c 1 -o gpurun_out/prof_r1l python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan > gpurun_out/ncu_l2.log 2>&1; tail -2 gpurun_out/ncu_l2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1l.json').read().strip().splitlines()[-1])
print(f"{d['value']:.4g} rays/s {d['ms_per_step']:.1f} ms e2e {d['e2e']['value']:.4g} ({d['e2e']['ms_per_step']:.1f} ms) frac {d['roofline']['frac']:.3f} prune {d['pruning']['gpu_ms']:.1f} ms x{d['pruning']['speedup']:.0f} same {d['pruning']['selected_set_identical']} cov {d['coverage_evaluation']['gpu_ms']:.2f} ms cpu {d['cpu_baseline']}")
print(open('gpurun_out/bench_r1l_ref.json').read()[:600])
PY
