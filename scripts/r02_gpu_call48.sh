#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "t5_encoder" > gpurun_out/r02au_pytest_t5.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/r02au_pytest_t5.log
timeout -k 10 300 python scripts/t5_bench.py > gpurun_out/r02au_t5_bench.txt 2>&1; echo rc=$?
tail -3 gpurun_out/r02au_t5_bench.txt
