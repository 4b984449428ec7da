#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02o_attn_sw_anatomy.txt
: > $out
for cfg in "SJD_ATTN=sw SJD_DEBUG_ATTN=1" "SJD_ATTN=sw SJD_BENCH_L=100" "SJD_ATTN=mma SJD_BENCH_L=100" "SJD_ATTN=sw SJD_BENCH_L=600" "SJD_ATTN=mma SJD_BENCH_L=600" "SJD_ATTN=sw SJD_BENCH_L=2400" "SJD_ATTN=mma SJD_BENCH_L=2400"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 32 2>&1 | grep "W=" >> $out
done
echo "== stamps sw W=32 L=1200" >> $out
SJD_ATTN=sw $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -9 >> $out
echo "== stamps sw W=32 L=100" >> $out
SJD_ATTN=sw $T 100 python scripts/attn_sw_stamps.py 32 100 2>&1 | tail -9 >> $out
cat $out
