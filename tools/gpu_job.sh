#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02l_gputest_multi_4gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02l_gputest_multi_4gpu.txt
tail -n 4 gpurun_out/r02l_gputest_multi_4gpu.txt
