cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_map.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python tools/diag/c3_pass.py 300 150000 4 2>&1 | grep -E "host wall|points|surfel_fit"
python bench.py --config C4 --steps 5 --warmup 3 > gpurun_out/r2z_c4_n1.json 2> gpurun_out/r2z_c4_n1.err; tail -c 300 gpurun_out/r2z_c4_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2z_c4_n1.json').read().strip().splitlines()[-1])
print('C4 N=1', d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, {k: round(v,3) for k,v in list(d.get('kernels_ms',{}).items())[:8]})
"
LVI_SURFEL_CLUSTER_FIT=0 python bench.py --config C4 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C4 N=1 no cluster', d['value']/1e9, d['ms_per_step'], {k: round(v,3) for k,v in list(d.get('kernels_ms',{}).items())[:4]})
"
