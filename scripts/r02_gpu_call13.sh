#!/bin/bash
# 8-GPU: headline bench + config 5 window sweep (accepted-len vs tokens/s at 8 GPUs)
mkdir -p gpurun_out
T="timeout -k 10"
$T 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r02m_bench_8gpu.json 2> gpurun_out/r02m_bench_8gpu.err; echo "8gpu bench rc=$?"
grep '^{' gpurun_out/r02m_bench_8gpu.json | cut -c1-500
$T 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --config 5 --steps 1 --warmup 1 > gpurun_out/r02m_bench_8gpu_config5.json 2> gpurun_out/r02m_bench_8gpu_config5.err; echo "8gpu config5 rc=$?"
grep '^{' gpurun_out/r02m_bench_8gpu_config5.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['n_gpus'])
for r in d['window_sweep']: print(r['window'], r['tokens_per_s'], r['accepted_tokens_per_iter'], r['ms_per_nfe'])
"
