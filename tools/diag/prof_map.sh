cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_map.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/r2z_c3_150k.json 2> gpurun_out/r2z_c3_150k.err
python bench.py --config C3 --leaves 5000000 --steps 5 --warmup 3 > gpurun_out/r2z_c3_5m.json 2> gpurun_out/r2z_c3_5m.err
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -c 300 gpurun_out/r2z_bench_n1.err
