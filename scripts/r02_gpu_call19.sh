#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02r_attn_sw.txt
: > $out
for cfg in "SJD_ATTN_SW_MERGE=0" "SJD_ATTN_SW_MERGE=1"; do
echo "== stamps sw W=32 L=1200 $cfg" >> $out
env SJD_ATTN=sw $cfg $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -16 | grep -v "unit [5-7]" >> $out
done
echo "== stamps sw W=32 L=100" >> $out
env SJD_ATTN=sw SJD_ATTN_SW_MERGE=0 $T 100 python scripts/attn_sw_stamps.py 32 100 2>&1 | tail -16 | grep -v "unit [1-7]" >> $out
cat $out
