#!/bin/bash
# Tuning builds: tools/variants.sh name "-DFLAG=.. ..." [name flags ...]  ->  build/variants/liblbm_b200_<name>.so (built in parallel).
# Run one with LBM_B200_LIB=build/variants/liblbm_b200_<name>.so python tools/f2_sweep.py ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags -shared \
      lbm-asynchronous_b200/csrc/lbm_b200.cu -o build/variants/liblbm_b200_$name.so -lcudart 2> build/variants/ptxas_$name.log \
      && echo "built $name" || (echo "FAILED $name"; tail -5 build/variants/ptxas_$name.log) ) &
done
wait
