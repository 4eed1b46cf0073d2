python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu --no-cov --no-plan > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['pruning'])"
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
