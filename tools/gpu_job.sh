#!/bin/bash
# scratch driver for one gpurun call (every step under its own timeout)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/p3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p3_pytest.log
tail -4 gpurun_out/p3_pytest.log
: > gpurun_out/p3_sweep.log
for k in 11621 208432 216831 216461 212441; do timeout 60 python tools/f2_sweep.py 8192 8192 64 $k >> gpurun_out/p3_sweep.log 2>&1 || echo "kernel $k rc=$?" >> gpurun_out/p3_sweep.log; done
for k in 10841::fast 208432::fast 216831::fast 212441::fast; do timeout 60 python tools/f2_sweep.py 8192 8192 64 $k >> gpurun_out/p3_sweep.log 2>&1 || echo "kernel $k rc=$?" >> gpurun_out/p3_sweep.log; done
timeout 60 python tools/f2_sweep.py 1024 1024 2000 0 >> gpurun_out/p3_sweep.log 2>&1
timeout 60 python tools/f2_sweep.py 32768 4096 32 208432 11621 >> gpurun_out/p3_sweep.log 2>&1
cat gpurun_out/p3_sweep.log
