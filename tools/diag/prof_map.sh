cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_map.py -m gpu -x -q 2>&1 | tail -3
python tools/diag/c3_pass.py 300 150000 4 2>&1 | tail -16
python tools/diag/c3_pass.py 300 5000000 4 2>&1 | tail -12
