#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=8000
rm -f gpurun_out/band_rows_sweep.txt
LBM_SWEEP_GPUS=1 timeout 300 python tools/small_sweep.py 10000 512x512,1024x512,2048x512,640x480,1024x1024 0 204 >> gpurun_out/band_rows_sweep.txt 2>&1
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 10000 1024x1024,2048x1024 0 204 >> gpurun_out/band_rows_sweep.txt 2>&1
cat gpurun_out/band_rows_sweep.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02i_gputest_multi_2gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_gputest_multi_2gpu.txt
tail -n 4 gpurun_out/r02i_gputest_multi_2gpu.txt
