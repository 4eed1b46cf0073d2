# usage: bash profiles/scripts/r2_bench_and_ncu.sh TAG   -> gpurun_out/TAG_*  (bench first, ncu only after it exited 0)
T=$1
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
FCVM_PROFILE=2 python profiles/scenes_timing.py christ > gpurun_out/${T}_christ_profile.jsonl 2> gpurun_out/${T}_christ_profile.err
python profiles/scenes_timing.py > gpurun_out/${T}_scenes_timing.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan --no-full --no-golden --no-stages > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none --import-source on -k regex:k_eval -s 3 -c 1 -o gpurun_out/${T}_k_eval python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan --no-full --no-golden --no-stages > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
