#!/bin/bash
# ncu captures for profiles/ (run under gpurun, one GPU).  Usage: tools/profile_gpu.sh <tag>
# 1. launch list of a short bench run (cold-cache, serialised: compare SHARES);  2. --set full of the kernels that matter.
set -u
TAG=${1:-r1}
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-calibration --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_factor_ll -s 4 -c 1 -f -o $OUT/prof_band_factor_$TAG $BENCH > $OUT/prof_band_factor_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linearize_kernel|build_system|band_backsolve|schur_eliminate' -s 12 -c 9 -f -o $OUT/prof_solver_$TAG $BENCH > $OUT/prof_solver_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'assoc_hit|assoc_select|voxel_leaf_stats|voxel_key|voxel_compact|undistort_kernel|surfel_fit' -c 8 -f -o $OUT/prof_map_$TAG $BENCH > $OUT/prof_map_$TAG.log 2>&1
ls -la $OUT | tail -12
