# usage: bash profiles/scripts/variant_los.sh default vA vB ...   (libraries under fc-planner_b200/_variants/): line-of-sight kernel ms for 2048 x 2048 pairs
cat > /tmp/planprof.py <<'PY'
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import bench
args = argparse.Namespace(los_points=2048)
d = bench.planner_grid_leg(args, 0, with_cpu=False)
print(sys.argv[1], 'los_kernel_ms', round(d['los_kernel_ms'], 3), 'los_gpu_ms', round(d['los_gpu_ms'], 2), 'lookups', d['los_voxel_lookups'], 'safe', d['los_safe_pairs'])
PY
for v in "$@"; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  python /tmp/planprof.py $v 2>/dev/null | tail -1
done
