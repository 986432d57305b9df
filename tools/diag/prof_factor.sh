cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:band_factor_ll -s 6 -c 1 -f -o gpurun_out/prof_factor_r2z python bench.py --steps 2 --warmup 3 --no-calibration --no-cpu-baseline > gpurun_out/prof_factor_r2z.log 2>&1
tail -3 gpurun_out/prof_factor_r2z.log
