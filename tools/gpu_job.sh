#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=8000
rm -f gpurun_out/band_multi_diag.txt
LBM_DEBUG=1 LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 2000 1024x1024 0 2>&1 | grep -v "graph replay\|dbg" | head -12 >> gpurun_out/band_multi_diag.txt
LBM_SWEEP_GPUS=1 timeout 300 python tools/small_sweep.py 10000 1024x512 204 514 >> gpurun_out/band_multi_diag.txt 2>&1
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 10000 1024x1024 514 204 >> gpurun_out/band_multi_diag.txt 2>&1
LBM_BAND_SM=0 LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 10000 1024x1024 514 >> gpurun_out/band_multi_diag.txt 2>&1
cat gpurun_out/band_multi_diag.txt
