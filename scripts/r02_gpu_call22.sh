#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and (chameleon or emu3) and sw" > gpurun_out/r02u_pytest_sw_toy.log 2>&1; rc=$?; echo "sw toy parity rc=$rc"
tail -3 gpurun_out/r02u_pytest_sw_toy.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
out=gpurun_out/r02u_attn_sw.txt
: > $out
for cfg in "SJD_ATTN=sw" "SJD_ATTN=sw SJD_BENCH_L=2400"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
echo "== stamps sw W=32 L=1200" >> $out
env SJD_ATTN=sw $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -16 | grep -v "unit [4-7]" >> $out
for cfg in "SJD_ATTN_SW_AUTO=0" "SJD_ATTN=sw"; do
  echo "== emu3 sweep $cfg" >> $out
  env $cfg $T 300 python scripts/config_sweep.py emu3 2>&1 | grep config >> $out
done
cat $out
