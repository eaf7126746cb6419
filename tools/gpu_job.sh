#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "give_up or per_process_api" > gpurun_out/timeout_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/timeout_tests.txt
tail -n 30 gpurun_out/timeout_tests.txt
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
