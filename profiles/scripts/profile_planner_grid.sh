cat > /tmp/planprof.py <<'PY'
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import bench
args = argparse.Namespace(los_points=2048)
print(bench.planner_grid_leg(args, 0, with_cpu=False))
PY
python /tmp/planprof.py | cut -c1-400
ncu --set full --clock-control none --import-source on -k regex:k_line_of_sight -s 1 -c 1 -o gpurun_out/prof_los_r1 python /tmp/planprof.py > gpurun_out/ncu_los.log 2>&1; tail -2 gpurun_out/ncu_los.log
