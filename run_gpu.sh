set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 2 --warmup 1 --views 4096 --no-cpu > gpurun_out/plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 1 --views 4096 --no-cpu > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_eval -s 1 -c 1 -o gpurun_out/prof_r1c python bench.py --steps 2 --warmup 1 --views 4096 --no-cpu > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log | cut -c1-200
