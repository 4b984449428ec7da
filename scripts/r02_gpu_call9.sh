#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python scripts/gemm_stamps2.py 32 > gpurun_out/r02i_stamps_w32.txt 2>&1; cat gpurun_out/r02i_stamps_w32.txt | tail -20
