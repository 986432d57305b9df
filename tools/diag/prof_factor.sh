cd $GRAFT_REPO_ROOT
run() { echo "== $1"; env $1 python bench.py --steps 20 --warmup 3 --no-calibration --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['phases_ms'], d['roofline']['kernels_ms'])"; }
run "X=1"
run "X=2"
run "LVI_PRE_SHIFT=3"
