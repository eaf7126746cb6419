#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_pytest.log
tail -n 4 gpurun_out/f2_pytest.log
timeout 200 python bench.py --arith fast --no-cpu-baseline > gpurun_out/f2_bench_n1_fast.json 2> gpurun_out/f2_bench_n1_fast.err; echo "bench fast rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/f2_bench_n1_fast.json < /dev/null
