#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02k_attn_tc_l2pf.txt
: > $out
for cfg in "SJD_ATTN=mma" "SJD_ATTN=tc SJD_ATTN_L2PF=1" "SJD_ATTN=tc SJD_ATTN_L2PF=0" "SJD_ATTN=tct SJD_ATTN_L2PF=1" "SJD_ATTN=tct SJD_ATTN_L2PF=0" "SJD_ATTN=mma SJD_ATTN_L2PF=0"; do
  echo "== $cfg" >> $out
  env $cfg $T 200 python scripts/chain_time.py 8 16,32,64 2>&1 | grep "W=" >> $out
done
cat $out
