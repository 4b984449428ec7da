#!/bin/bash
# 2-GPU validation: bench under torchrun (ours + reference arm), multi-GPU prompt runner
mkdir -p gpurun_out
T="timeout -k 10"
$T 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02j_bench_2gpu.json 2> gpurun_out/r02j_bench_2gpu.err; echo "2gpu bench rc=$?"
grep '^{' gpurun_out/r02j_bench_2gpu.json | cut -c1-700; tail -3 gpurun_out/r02j_bench_2gpu.err | cut -c1-300
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-budget 6 > gpurun_out/r02j_bench_2gpu_reference.json 2>/dev/null; echo "2gpu reference arm rc=$?"
grep '^{' gpurun_out/r02j_bench_2gpu_reference.json | cut -c1-300
rm -rf /tmp/sjd_workdir
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 scripts/run_prompts.py --model synthetic/lumina-mgpt-7b-768 --n_layers 4 --target_size 256 --output-dir /tmp/sjd_workdir > gpurun_out/r02j_launcher.log 2>&1; echo "launcher rc=$?"
grep launcher gpurun_out/r02j_launcher.log; ls /tmp/sjd_workdir | tr '\n' ' '
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --config 1 --steps 2 --warmup 1 > gpurun_out/r02j_bench_2gpu_config1.json 2>/dev/null; echo "2gpu config1 rc=$?"
grep '^{' gpurun_out/r02j_bench_2gpu_config1.json | cut -c1-300
