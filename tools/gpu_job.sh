#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=5000
timeout 300 python tools/small_sweep.py 10000 1024x1024,1024x512,2048x1024 514 508 > gpurun_out/band64_sweep.txt 2>&1
cat gpurun_out/band64_sweep.txt
