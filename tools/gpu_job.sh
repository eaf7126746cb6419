#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02d_gputest_multi_4gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_gputest_multi_4gpu.txt
tail -n 4 gpurun_out/r02d_gputest_multi_4gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 100 --warmup 10 > gpurun_out/r02d_bench_n4.json 2> gpurun_out/r02d_bench_n4.err; echo "bench rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02d_bench_n4.json < /dev/null
tail -n 3 gpurun_out/r02d_bench_n4.err
