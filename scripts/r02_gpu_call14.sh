#!/bin/bash
# attention_sw.cu: first parity + timing
mkdir -p gpurun_out
T="timeout -k 10"
$T 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and chameleon and sw" > gpurun_out/r02n_pytest_sw_toy.log 2>&1; rc=$?; echo "sw toy parity rc=$rc"
tail -5 gpurun_out/r02n_pytest_sw_toy.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "window_forward and (chameleon or emu3) and sw" > gpurun_out/r02n_pytest_sw_toy_all.log 2>&1; echo "sw toy parity (all variants) rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r02n_pytest_sw_toy_all.log | tail -12
out=gpurun_out/r02n_attn_sw.txt
: > $out
for cfg in "SJD_ATTN=mma" "SJD_ATTN=sw" "SJD_ATTN=sw SJD_ATTN_SW_NCOLS=64" "SJD_ATTN=tc" "SJD_ATTN=sw SJD_ATTN_SW_GROW=0"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
echo "== stamps sw W=32" >> $out
SJD_ATTN=sw $T 100 python scripts/attn_stamps.py 32 >> $out 2>&1
tail -9 $out
$T 500 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "full_width and sw" > gpurun_out/r02n_pytest_sw_full.log 2>&1; echo "sw full-width parity rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/r02n_pytest_sw_full.log | tail -12
