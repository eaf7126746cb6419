#!/bin/bash
# scratch driver for one gpurun call (every step under its own timeout)
timeout 300 python -m pytest tests/test_gpu_fused2.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/p9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p9_pytest.log
tail -n 3 gpurun_out/p9_pytest.log
: > gpurun_out/p9_sweep.log
for k in 208432 11621 10841 208432::fast 10841::fast; do timeout 60 python tools/f2_sweep.py 8192 8192 64 $k >> gpurun_out/p9_sweep.log 2>&1 || echo "kernel $k rc=$?" >> gpurun_out/p9_sweep.log; done
timeout 60 python tools/f2_sweep.py 1024 1024 2000 0 >> gpurun_out/p9_sweep.log 2>&1
cat gpurun_out/p9_sweep.log
