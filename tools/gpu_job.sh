#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_gputest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_gputest.txt
tail -n 4 gpurun_out/r02n_gputest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
