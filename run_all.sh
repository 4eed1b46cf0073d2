timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
FCVM_PROFILE=1 python profiles/fullscale_update.py 25000 2>&1 | tail -4
