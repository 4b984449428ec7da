#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02s_attn_sw.txt
: > $out
for L in 100 1200; do
echo "== stamps sw W=32 L=$L" >> $out
env SJD_ATTN=sw SJD_ATTN_SW_MERGE=0 $T 100 python scripts/attn_sw_stamps.py 32 $L 2>&1 | tail -16 | grep -v "unit [4-7]" >> $out
done
cat $out
