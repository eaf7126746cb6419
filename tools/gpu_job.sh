#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_gputest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_gputest.txt
tail -n 4 gpurun_out/r02j_gputest.txt
timeout 900 python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; echo "bench rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02j_bench_n1.json < /dev/null
tail -n 3 gpurun_out/r02j_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
