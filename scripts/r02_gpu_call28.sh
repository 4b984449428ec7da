#!/bin/bash
mkdir -p gpurun_out
SJD_STAMPS_CLASSES=1 timeout -k 10 100 python scripts/attn_sw_stamps.py 32 1200 2>&1 | tail -22 > gpurun_out/r02aa_attn_sw_classes.txt
cat gpurun_out/r02aa_attn_sw_classes.txt
