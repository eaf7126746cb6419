#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
export LBM_HALO_TIMEOUT_MS=8000
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 0 1024x1024 0 11041 > gpurun_out/loop2_sweep.txt 2>&1
LBM_SWEEP_GPUS=2 timeout 300 python tools/small_sweep.py 5000 2048x512,1024x512,4096x256 0 11041 >> gpurun_out/loop2_sweep.txt 2>&1
cat gpurun_out/loop2_sweep.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02e_gputest_multi.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_gputest_multi.txt
tail -n 4 gpurun_out/r02e_gputest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err; echo "bench rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02e_bench_n2.json < /dev/null
tail -n 3 gpurun_out/r02e_bench_n2.err
