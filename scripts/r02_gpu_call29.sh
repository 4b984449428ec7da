#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and (chameleon or emu3) and sw" > gpurun_out/r02ab_pytest_sw_toy.log 2>&1; rc=$?; echo "sw toy parity rc=$rc"
tail -3 gpurun_out/r02ab_pytest_sw_toy.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
out=gpurun_out/r02ab_attn_sw.txt
: > $out
for cfg in "SJD_ATTN_SW_RUNCOST=0" "SJD_ATTN_SW_RUNCOST=0.85" "SJD_ATTN_SW_RUNCOST=0.5" "SJD_ATTN_SW_RUNCOST=1.2"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" | sed 's/gemm-only.*| //' >> $out
done
for cfg in "SJD_ATTN_SW_RUNCOST=0" "SJD_ATTN_SW_RUNCOST=0.85"; do
echo "== stamps $cfg" >> $out
env $cfg SJD_STAMPS_CLASSES=1 $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -22 | grep -v "^unit\|producer" >> $out
done
cat $out
