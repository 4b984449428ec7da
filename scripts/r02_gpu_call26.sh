#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python scripts/gemm_stamps3.py 32 > gpurun_out/r02y_chain_stragglers.txt 2>&1
cat gpurun_out/r02y_chain_stragglers.txt | tail -12
