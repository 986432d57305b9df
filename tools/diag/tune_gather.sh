#!/bin/bash
# variants of the gather kernel: rebuild assemble.cu.o with different constants and time one iteration
set -u
cd $GRAFT_REPO_ROOT
NVCC=/usr/local/cuda/bin/nvcc
build() {  # $1 = extra defines
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -Xcudafe --diag_suppress=177 $1 -c lvi_exc_b200/csrc/assemble.cu -o build/assemble.cu.o 2>&1 | grep -v warning | head -5
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o lvi_exc_b200/lib/liblvi_exc_b200.so build/*.cu.o -lcudart -ldl
}
run() { echo "== $1 ctas=$2"; build "$1"; LVI_GATHER_CTAS=$2 python tools/diag/linearize_times.py 2>&1 | grep -E "phases|gather_kernel " ; }
run "" 3
run "-DLVI_STAGE_DOUBLES=1024" 5
run "-DLVI_STAGE_DOUBLES=1024 -DLVI_GATHER_WARPS=2" 10
run "-DLVI_STAGE_DOUBLES=1280" 4
run "-DLVI_GATHER_CHUNK=192" 3
run "-DLVI_GATHER_CHUNK=48" 3
run "-DLVI_STAGE_DOUBLES=1024 -DLVI_GATHER_CHUNK=192" 5
