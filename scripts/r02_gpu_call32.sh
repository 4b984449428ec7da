#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02ae_epi4.txt
: > $out
for cfg in "SJD_GEMM_EPI2=1" "SJD_GEMM_EPI2=4" "SJD_GEMM_EPI2=1" "SJD_GEMM_EPI2=4"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
SJD_GEMM_EPI2=4 $T 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and auto or real_stack or do_not_depend or gemm" > gpurun_out/r02ae_pytest.log 2>&1; echo "pytest EPI2=4 rc=$?"
tail -3 gpurun_out/r02ae_pytest.log
