#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=8000
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02m_gputest_multi_4gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_gputest_multi_4gpu.txt
tail -n 4 gpurun_out/r02m_gputest_multi_4gpu.txt
rm -f gpurun_out/ll2_multi_sweep.txt
LBM_SWEEP_GPUS=4 timeout 200 python tools/small_sweep.py 0 1024x1024 0 404 0::fast >> gpurun_out/ll2_multi_sweep.txt 2>&1
LBM_SWEEP_GPUS=1 timeout 200 python tools/small_sweep.py 0 1024x1024 0 >> gpurun_out/ll2_multi_sweep.txt 2>&1
cat gpurun_out/ll2_multi_sweep.txt
