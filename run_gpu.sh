set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_r1c.json; tail -5 gpurun_out/bench_r1c.err
