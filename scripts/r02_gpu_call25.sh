#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02x_attn_crossover.txt
: > $out
for L in 64 200 400 600 800; do
for cfg in "SJD_ATTN=mma" "SJD_ATTN=sw"; do
  echo "== $cfg L=$L" >> $out
  env $cfg SJD_BENCH_L=$L $T 150 python scripts/chain_time.py 8 16,32 2>&1 | grep "W=" | sed 's/gemm-only.*| //' >> $out
done
done
cat $out
$T 1500 python -m pytest tests -q -m gpu > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02x_pytest_gpu.log
