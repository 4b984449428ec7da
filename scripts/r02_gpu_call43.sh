#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "long_cache" > gpurun_out/r02ap_pytest_long.log 2>&1; echo "rc=$?"
tail -15 gpurun_out/r02ap_pytest_long.log
