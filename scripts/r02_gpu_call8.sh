#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02h_rowc8.txt
: > $out
for cfg in "SJD_GEMM_ROWC8=0" "SJD_GEMM_ROWC8=1" "SJD_GEMM_ROWC8=0" "SJD_GEMM_ROWC8=1"; do
  echo "== $cfg" >> $out
  env $cfg $T 200 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
SJD_GEMM_ROWC8=1 $T 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -k "gemm or window_forward or full_width or real_stack or end_to_end" > gpurun_out/r02h_pytest.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r02h_pytest.log | cut -c1-300
