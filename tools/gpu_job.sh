#!/bin/bash
# scratch job script for gpurun (rewritten per call)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err; echo "bench rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02k_bench_n1.json < /dev/null
tail -n 2 gpurun_out/r02k_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_ref_n1.json 2> gpurun_out/r02k_bench_ref_n1.err; echo "ref rc=$?"
tail -c 600 gpurun_out/r02k_bench_ref_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02k_bench_n2.json 2> gpurun_out/r02k_bench_n2.err; echo "bench2 rc=$?"
timeout 20 python tools/bench_line.py gpurun_out/r02k_bench_n2.json < /dev/null
tail -n 2 gpurun_out/r02k_bench_n2.err
