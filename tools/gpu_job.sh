#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_gputest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_gputest.txt
tail -n 4 gpurun_out/r02f_gputest.txt
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02f_bench_n1.json < /dev/null
tail -n 3 gpurun_out/r02f_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_steps40.csv python bench.py --steps 40 --warmup 3 --no-shipped > gpurun_out/r02f_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_band -c 1 -o gpurun_out/band_514 -f python tools/small_sweep.py 300 1024x1024 514 > gpurun_out/ncu_band2.log 2>&1
tail -2 gpurun_out/ncu_band2.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
