#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=5000
rm -f gpurun_out/ll_multi_sweep.txt
LBM_SWEEP_GPUS=1 timeout 300 python tools/small_sweep.py 0 1024x1024 0 >> gpurun_out/ll_multi_sweep.txt 2>&1
LBM_SWEEP_GPUS=4 timeout 300 python tools/small_sweep.py 0 1024x1024 0 404 11041 204 404::fast 11041::fast >> gpurun_out/ll_multi_sweep.txt 2>&1
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 0 1024x1024 0 11041 204 >> gpurun_out/ll_multi_sweep.txt 2>&1
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 0 128x128,128x256,256x256 0 404 11041 201 >> gpurun_out/ll_multi_sweep.txt 2>&1
LBM_SWEEP_GPUS=4 timeout 300 python tools/small_sweep.py 0 256x256 0 404 11041 >> gpurun_out/ll_multi_sweep.txt 2>&1
cat gpurun_out/ll_multi_sweep.txt
