#!/bin/bash
# usage: scale_c4.sh N tag -- torchrun of the C4 map-path bench at N GPUs
N=$1; TAG=$2
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config C4 --steps 5 --warmup 3 > gpurun_out/${TAG}_c4_n$N.json 2> gpurun_out/${TAG}_c4_n$N.err
python - <<P
import json
d=json.loads(open("gpurun_out/${TAG}_c4_n$N.json").read().strip().splitlines()[-1])
print("C4 N=$N", d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, {k: round(v,3) for k,v in list(d.get("kernels_ms",{}).items())[:8]})
P
