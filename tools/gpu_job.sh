#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "one_process_per_gpu" > gpurun_out/r02o_perprocess.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_perprocess.txt
tail -n 4 gpurun_out/r02o_perprocess.txt
