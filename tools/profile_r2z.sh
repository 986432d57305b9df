#!/bin/bash
# end-of-round captures (one GPU): test log, full bench line, ncu launch list of the short bench, --set full of the factor / gather kernels, C3
OUT=gpurun_out
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -q > $OUT/r2z_gpu_tests.log 2>&1; tail -2 $OUT/r2z_gpu_tests.log
python bench.py > $OUT/r2z_bench_n1.json 2> $OUT/r2z_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r2z_reference_arm.json 2> $OUT/r2z_reference_arm.err
BENCH="python bench.py --steps 2 --warmup 3 --no-calibration --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2z_launches_bench_steps2.csv $BENCH > $OUT/r2z_launches.bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'band_factor_ll|gather_kernel' -s 4 -c 3 -f -o $OUT/r2z_prof_solver $BENCH > $OUT/r2z_prof_solver.log 2>&1
python bench.py --config C3 --steps 5 --warmup 3 > $OUT/r2z_c3_150k.json 2> $OUT/r2z_c3_150k.err
python bench.py --config C3 --leaves 5000000 --steps 5 --warmup 3 > $OUT/r2z_c3_5m.json 2> $OUT/r2z_c3_5m.err
tail -c 400 $OUT/r2z_bench_n1.json
