set -x
nvidia-smi -L | head -8
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --steps 3 --warmup 3 --no-cpu --no-cov --no-plan > gpurun_out/scale_r1b_n$n.json 2> gpurun_out/scale_r1b_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 --no-cov --no-plan > gpurun_out/scale_r1b_n$n.json 2> gpurun_out/scale_r1b_n$n.err; fi
  echo "n=$n rc=$?"; tail -1 gpurun_out/scale_r1b_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('pruning'))"
  tail -3 gpurun_out/scale_r1b_n$n.err
done
