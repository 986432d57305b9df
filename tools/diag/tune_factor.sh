#!/bin/bash
# variants of the factor kernel: rebuild solver.cu.o with different constants and time the phases
set -u
cd $GRAFT_REPO_ROOT
NVCC=/usr/local/cuda/bin/nvcc
build() {
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Xcudafe --diag_suppress=177 $1 -c lvi_exc_b200/csrc/solver.cu -o build/solver.cu.o 2>&1 | grep -E "error" | head -5
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o lvi_exc_b200/lib/liblvi_exc_b200.so build/*.cu.o -lcudart -ldl
}
run() { echo "== $1 $2"; build "$1"; env $2 python bench.py --steps 20 --warmup 3 --no-calibration --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['phases_ms']['band_factor'], d['phases_ms']['corner_backsolve'], d['e2e']['final_cost'])"; }
for v in "$@"; do run "$v" "X=1"; done
