#!/bin/bash
# round 2, GPU call 3: Philox transform pinning, lookahead / KV-prefetch sweeps at tpu 1, suites, bench
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "philox" > gpurun_out/r02c_pytest_philox.log 2>&1; rc=$?
echo "philox rc=$rc"; grep -E "mismatches per variant|passed|failed" gpurun_out/r02c_pytest_philox.log | cut -c1-400 | head -12
if [ $rc -ne 0 ]; then export SJD_HOST_NOISE=1; echo "PHILOX MISMATCH -> SJD_HOST_NOISE=1 for the rest"; fi
out=gpurun_out/r02c_chain_experiments.txt
: > $out
for cfg in "SJD_GEMM_LOOKAHEAD=32" "SJD_GEMM_LOOKAHEAD=0" "SJD_GEMM_LOOKAHEAD=8" "SJD_GEMM_LOOKAHEAD=16" "SJD_GEMM_LOOKAHEAD=64" "SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=8" "SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=16" "SJD_KV_PF=0" "SJD_KV_PF=0 SJD_ATTN_L2PF=0" "SJD_ATTN_L2PF=0"; do
  echo "== $cfg" >> $out
  env $cfg $T 200 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
$T 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "gpu suites rc=$?"
tail -8 gpurun_out/r02c_pytest_gpu.log
$T 200 python scripts/gemm_stamps.py 32 > gpurun_out/r02c_gemm_stamps_w32.txt 2>&1
$T 800 python bench.py --steps 2 --warmup 3 --cpu-budget 8 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
cat gpurun_out/r02c_bench.json | cut -c1-1800; tail -2 gpurun_out/r02c_bench.err
