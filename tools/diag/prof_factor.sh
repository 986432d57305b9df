cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -x -q 2>&1 | tail -3
timeout 600 bash tools/diag/tune_factor.sh "-DLVI_FAC_STAGES=3" "-DLVI_FAC_STAGES=4" "-DLVI_FAC_STAGES=5"
