# launch list (per-kernel times) of the planner-grid leg and a --set full capture of k_los_march, round 2
cat > /tmp/planprof.py <<'PY'
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import bench
args = argparse.Namespace(los_points=2048)
print(bench.planner_grid_leg(args, 0, with_cpu=False))
PY
python /tmp/planprof.py | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${1:-r2y}_launches_plan.csv python /tmp/planprof.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_los_march -s 1 -c 1 -o gpurun_out/${1:-r2y}_los python /tmp/planprof.py > gpurun_out/ncu_los.log 2>&1; tail -2 gpurun_out/ncu_los.log
