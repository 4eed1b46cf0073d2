python -m pytest tests/test_multi_gpu.py tests/test_cpp_shim.py -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu --no-cov --no-plan > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['pruning'])"
