#!/bin/bash
mkdir -p gpurun_out
env SJD_ATTN=sw timeout -k 10 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -14 | grep -v "unit [0-7]" > gpurun_out/r02az_cluster_tail.txt
cat gpurun_out/r02az_cluster_tail.txt
