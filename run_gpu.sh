python -m pytest tests/test_gpu_bounds.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for st in 3 4; do python bench.py --stage $st --steps 3 --warmup 2 --no-prune --cpu-seconds 6 > gpurun_out/bench_r1g_E$st.json 2> gpurun_out/bench_r1g_E$st.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1g_E$st.json').read().strip().splitlines()[-1])
print('E$st', f"{d['value']:.4g} rays/s {d['ms_per_step']:.1f} ms e2e {d['e2e']['value']:.4g}", d['rays_per_step_per_gpu'], d['voxel_lookups_per_step_per_gpu'], d['roofline']['frac'], d['cpu_baseline'])
PY
done
FCVM_PROFILE=1 python profiles/fullscale_update.py 25000 > gpurun_out/fullscale_update.json 2> gpurun_out/fullscale_update.err; cat gpurun_out/fullscale_update.json; tail -3 gpurun_out/fullscale_update.err
