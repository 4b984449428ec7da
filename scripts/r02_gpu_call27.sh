#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "window_forward and auto or real_stack or do_not_depend or gemm" > gpurun_out/r02z_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -3 gpurun_out/r02z_pytest.log
if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "HANG: stopping"; exit 0; fi
out=gpurun_out/r02z_weighted_streamk.txt
: > $out
for cfg in "SJD_SK_WEIGHTED=0" "SJD_SK_WEIGHTED=1" "SJD_SK_E=2,8,10" "SJD_SK_E=3,16,18" "SJD_SK_E=3,20,22" "SJD_SK_E=1,12,12"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
echo "== stragglers weighted" >> $out
$T 100 python scripts/gemm_stamps3.py 32 2>&1 | grep "op E-t0" | cut -c1-220 >> $out
tail -5 $out
