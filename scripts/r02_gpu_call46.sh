#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "t5_encoder" > gpurun_out/r02as_pytest_t5.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/r02as_pytest_t5.log
