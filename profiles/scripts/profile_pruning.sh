FCVM_PROFILE=1 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import bench, argparse
args = argparse.Namespace(prune_points=40000, prune_views=4000)
out = bench.pruning_leg(args, 0, with_cpu=False)
print(out)
PY
