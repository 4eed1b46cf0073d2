# 2 / 4 / 8-GPU legs of the scaling sweep on one 8-GPU box (n = 1: profiles/scripts/final_bench.sh)
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cpu --no-cov --no-plan > gpurun_out/scale_r1d_n$n.json 2> gpurun_out/scale_r1d_n$n.err
  echo "n=$n rc=$?"; tail -1 gpurun_out/scale_r1d_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['pruning']['gpu_ms'], d['pruning'].get('ranks_agree'))"
done
