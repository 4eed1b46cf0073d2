python -m pytest tests/test_gpu_parity.py -x -q -s -k "lattice or evaluator" 2>&1 | tail -8
