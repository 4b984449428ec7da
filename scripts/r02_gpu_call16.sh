#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02o_attn_sw_anatomy2.txt
: > $out
for L in 100 1200; do
for d in 0 2 3 4; do
  echo "== sw L=$L SJD_DEBUG_ATTN=$d" >> $out
  SJD_ATTN=sw SJD_BENCH_L=$L SJD_DEBUG_ATTN=$d $T 150 python scripts/chain_time.py 8 32 2>&1 | grep "W=" >> $out
done
done
cat $out
