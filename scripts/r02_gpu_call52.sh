#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "long_cache" > gpurun_out/r02ay_pytest_long.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/r02ay_pytest_long.log
