#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and (chameleon or emu3) and (sw or auto)" > gpurun_out/r02bd_pytest_sw_toy.log 2>&1; rc=$?; echo "sw toy parity rc=$rc"
tail -3 gpurun_out/r02bd_pytest_sw_toy.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
out=gpurun_out/r02bd_narrow.txt
: > $out
for cfg in "SJD_ATTN_SW_CLUSTER_MINR=32" "SJD_ATTN_SW_CLUSTER_MINR=8"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 1,8,16,32 2>&1 | grep "W=" | sed 's/gemm-only.*| //' >> $out
done
cat $out
$T 900 python -m pytest tests -q -m gpu > gpurun_out/r02bd_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02bd_pytest_gpu.log
