python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --no-cov --no-plan > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo rc=$?
python -c "
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['pruning']['gpu_ms'], d['pruning']['ranks_agree'])"
