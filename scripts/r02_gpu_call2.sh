#!/bin/bash
# round 2, GPU call 2: tiles-per-unit GEMM correctness + timing, device Philox parity, then the suites and the bench.
mkdir -p gpurun_out
T="timeout -k 10"
$T 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or philox" > gpurun_out/r02b_pytest_gemm_philox.log 2>&1; rc=$?
echo "gemm+philox rc=$rc"; tail -15 gpurun_out/r02b_pytest_gemm_philox.log
if [ $rc -ne 0 ]; then
  $T 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm" > gpurun_out/r02b_pytest_gemm_only.log 2>&1; rcg=$?
  echo "gemm only rc=$rcg"; tail -5 gpurun_out/r02b_pytest_gemm_only.log
  if [ $rcg -ne 0 ]; then export SJD_GEMM_TPU=1; echo "FALLING BACK TO SJD_GEMM_TPU=1"; fi
fi
out=gpurun_out/r02b_chain_experiments.txt
: > $out
for cfg in "SJD_GEMM_TPU=2" "SJD_GEMM_TPU=1" "SJD_GEMM_TPU=4" "SJD_GEMM_TPU=2 SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=8" "SJD_GEMM_TPU=2 SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=24" "SJD_GEMM_TPU=2 SJD_GEMM_LOOKAHEAD=0" "SJD_GEMM_TPU=4 SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=12"; do
  echo "== $cfg" >> $out
  Ws=8,16,32,64,128; case "$cfg" in *TPU=4*) Ws=8,16,32,64;; esac
  env $cfg $T 240 python scripts/chain_time.py 8 $Ws 2>&1 | grep "W=" >> $out
done
cat $out
$T 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu > gpurun_out/r02b_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -8 gpurun_out/r02b_pytest_new.log
$T 900 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -8 gpurun_out/r02b_pytest_gpu.log
$T 200 python scripts/gemm_stamps.py 32 > gpurun_out/r02b_gemm_stamps_w32.txt 2>&1; head -8 gpurun_out/r02b_gemm_stamps_w32.txt
$T 800 python bench.py --steps 2 --warmup 3 --cpu-budget 8 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
cat gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
