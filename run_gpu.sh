set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_r1b.json; tail -5 gpurun_out/bench_r1b.err
python bench.py --steps 2 --warmup 1 --views 2048 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o gpurun_out/prof_r1a python bench.py --steps 2 --warmup 1 --views 2048 --no-cpu > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-300
