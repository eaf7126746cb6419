#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=5000
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "strict_steps_bit_exact or av_vels_identical or chunked" > gpurun_out/cl_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/cl_tests.txt
tail -n 3 gpurun_out/cl_tests.txt
timeout 300 python tools/small_sweep.py 0 128x128,128x256,256x256 401 404 401::fast > gpurun_out/ll_wait_sweep.txt 2>&1
timeout 300 python tools/small_sweep.py 20000 1024x256 404 >> gpurun_out/ll_wait_sweep.txt 2>&1
cat gpurun_out/ll_wait_sweep.txt
