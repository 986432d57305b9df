#!/bin/bash
OUT=gpurun_out
python bench.py > $OUT/r2m_bench_c2.json 2> $OUT/r2m_bench_c2.err
BENCH="python bench.py --steps 2 --warmup 3 --no-calibration --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2m_launches.csv $BENCH > $OUT/r2m_launches.bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|jacobian_kernel' -s 8 -c 5 -f -o $OUT/r2m_prof_assemble $BENCH > $OUT/r2m_prof_assemble.log 2>&1
tail -3 $OUT/r2m_bench_c2.json | cut -c1-600
