#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02bc_cluster_minr.txt
: > $out
for cfg in "SJD_ATTN_SW_CLUSTER_MINR=32" "SJD_ATTN_SW_CLUSTER_MINR=8" "SJD_ATTN_SW_CLUSTER_MINR=32 SJD_BENCH_L=300" "SJD_ATTN_SW_CLUSTER_MINR=8 SJD_BENCH_L=300"; do
  echo "== $cfg" >> $out
  env $cfg $T 150 python scripts/chain_time.py 8 8,16 2>&1 | grep "W=" | sed 's/gemm-only.*| //' >> $out
done
cat $out
