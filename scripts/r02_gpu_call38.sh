#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02ak_cluster_stamps.txt
: > $out
for cfg in "SJD_ATTN_SW_CLUSTER=0" "SJD_ATTN_SW_CLUSTER=4"; do
echo "== $cfg" >> $out
env SJD_ATTN=sw $cfg $T 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -12 | grep -v "unit [4-7]" >> $out
done
cat $out
