#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
D=accelerating-t2i-ar-with-sjd_b200
out=gpurun_out/r02ag_epi_warps.txt
: > $out
cp $D/libsjd_b200.so $D/libsjd_b200_keep.so
for v in "base SJD_GEMM_EPI2=1" "w12 SJD_GEMM_EPI2=1" "w12 SJD_GEMM_EPI2=4" "w12 SJD_GEMM_EPI2=0" "base SJD_GEMM_EPI2=1" "w12 SJD_GEMM_EPI2=4"; do
  set -- $v
  cp $D/libsjd_b200_$1.so $D/libsjd_b200.so
  echo "== lib $1 $2" >> $out
  env $2 $T 150 python scripts/chain_time.py 8 16,32,64,128 2>&1 | grep "W=" >> $out
done
cp $D/libsjd_b200_keep.so $D/libsjd_b200.so
cat $out
