#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and (chameleon or emu3) and sw" > gpurun_out/r02ba_pytest_sw_toy.log 2>&1; rc=$?; echo "sw toy parity rc=$rc"
tail -3 gpurun_out/r02ba_pytest_sw_toy.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
out=gpurun_out/r02ba_attn_sw_tail2.txt
: > $out
for cfg in "SJD_ATTN_SW_CLUSTER=0" "SJD_ATTN_SW_CLUSTER=4" "SJD_ATTN_SW_CLUSTER=0" "SJD_ATTN_SW_CLUSTER=4"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 32,64 2>&1 | grep "W=" | sed 's/gemm-only.*| //' >> $out
done
echo "== stamps" >> $out
env SJD_ATTN=sw $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -14 | grep -v "unit [0-7]" >> $out
cat $out
$T 600 python -m pytest tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -q -m gpu -k "full_width or do_not_depend or long_cache or real_stack or end_to_end" > gpurun_out/r02ba_pytest_sw_full.log 2>&1; echo "forward parity rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/r02ba_pytest_sw_full.log | tail -12
