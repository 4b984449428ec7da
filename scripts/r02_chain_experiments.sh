#!/bin/bash
# Round-2 GEMM-chain experiments (GPU box): is the chain L2-bandwidth bound by the activation-tile re-reads?
#   SJD_DEBUG_XSKIP=1     skip the X-tile loads (garbage results, timing only)
#   SJD_GEMM_PF_ALWAYS=1  keep the L2 prefetch frontier ahead in steady state
out=gpurun_out/r02_chain_experiments.txt
: > $out
for cfg in "" "SJD_DEBUG_XSKIP=1" "SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=8" "SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=16" "SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=32" "SJD_DEBUG_XSKIP=1 SJD_GEMM_PF_ALWAYS=1 SJD_GEMM_LOOKAHEAD=16"; do
  echo "== $cfg" >> $out
  env $cfg python scripts/chain_time.py 8 8,16,32,64,128 2>&1 | grep "W=" >> $out
done
cat $out
