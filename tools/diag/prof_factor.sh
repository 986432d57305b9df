#!/bin/bash
# factor-kernel check in one GPU call (diagnostics): solver parity tests, bench phases, critical-path trace
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_solver.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-calibration --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['phases_ms'], d['e2e']['value'], d['e2e']['final_cost'])"
LVI_TRACE_FACTOR=gpurun_out/trace.bin timeout 120 python bench.py --steps 5 --warmup 3 --no-calibration --no-cpu-baseline >/dev/null 2>&1
python tools/analyze_factor_trace.py gpurun_out/trace.bin 2>&1 | grep -v Warn | head -9
python tools/trace_tasks.py gpurun_out/trace.bin 2>&1 | grep -v "Warn\|nanq" | head -8
rm -f gpurun_out/trace.bin
