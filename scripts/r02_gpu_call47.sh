#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/t5_bench.py > gpurun_out/r02at_t5_bench.txt 2>&1; echo rc=$?
tail -3 gpurun_out/r02at_t5_bench.txt
