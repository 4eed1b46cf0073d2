for v in I J K L M; do
  export FCVM_LIB=/root/repo/fc-planner_b200/_variants/$v.so
  python bench.py --steps 3 --warmup 2 --no-cpu --no-prune > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1])
    print('$v', f"{d['ms_per_step']:.2f} ms/step", f"kernel {d['roofline']['kernel_ms']:.2f} ms", f"e2e {d['e2e']['ms_per_step']:.1f} ms")
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/var_$v.err').read()[-800:])
PY
done
unset FCVM_LIB
bash run_prof_prune.sh 2>&1 | tail -8
