#!/bin/bash
# 2-GPU sanity check of the final code (replicas; NCCL counter all-gather)
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02an_bench_2gpu.json 2> gpurun_out/r02an_bench_2gpu.err; echo "2gpu bench rc=$?"
grep '^{' gpurun_out/r02an_bench_2gpu.json | cut -c1-400; tail -3 gpurun_out/r02an_bench_2gpu.err
