# usage: bash profiles/scripts/variant_cov.sh default vA vB ...   (libraries under fc-planner_b200/_variants/): CoverageEvaluation of the MBS trajectory
cat > /tmp/covvar.py <<'PY'
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import bench
d = bench.coverage_leg(argparse.Namespace(), 0, with_cpu=False)
print(sys.argv[1], 'scan_ms', round(d['scan_kernel_ms'], 3), 'kernels_ms', round(d['kernel_ms'], 3), 'call_ms', round(d['gpu_ms'], 2), 'chunks', d['chunks_scanned'], 'fixture', d['matches_fixture'])
PY
for v in "$@"; do
  if [ $v = default ]; then unset FCVM_LIB; else export FCVM_LIB=$PWD/fc-planner_b200/_variants/$v.so; fi
  python /tmp/covvar.py $v 2>/dev/null | tail -1
done
