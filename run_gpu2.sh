set -x
python -m pytest tests/test_multi_gpu.py tests/test_cpp_shim.py -m gpu -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r1k_n2.json 2> gpurun_out/bench_r1k_n2.err; echo rc=$?
cut -c1-1500 gpurun_out/bench_r1k_n2.json; tail -5 gpurun_out/bench_r1k_n2.err
