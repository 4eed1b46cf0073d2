# usage: bash profiles/scripts/r2_final.sh TAG   -> gpurun_out/TAG_*  (what the driver runs at round end, then the ncu passes; ncu only after the bench exited 0)
T=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || exit 1
python profiles/scenes_timing.py > gpurun_out/${T}_scenes_timing.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan --no-full --no-golden --no-stages > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum --clock-control none --import-source on -k regex:k_eval -s 3 -c 1 -o gpurun_out/${T}_k_eval python bench.py --steps 2 --warmup 1 --no-cpu --no-prune --no-cov --no-plan --no-full --no-golden --no-stages > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
