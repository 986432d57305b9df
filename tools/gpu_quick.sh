#!/bin/bash
# solver tests, a short bench (phases) and a factor trace in one GPU call (diagnostics)
timeout 120 python -m pytest tests/test_gpu_solver.py -x -q 2>&1 | tail -3
timeout 120 python bench.py --steps 20 --warmup 3 --no-calibration --no-cpu-baseline 2>gpurun_out/b.err | tail -1 > gpurun_out/b.json
python -c "import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['phases_ms'], d['e2e']['value'])"
LVI_TRACE_FACTOR=gpurun_out/trace.bin timeout 120 python bench.py --steps 5 --warmup 3 --no-calibration --no-cpu-baseline >/dev/null 2>&1
python tools/analyze_factor_trace.py gpurun_out/trace.bin
