#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_baseline_sizes.py -q -m gpu -k "vq_decoder" > gpurun_out/r02ar_pytest_vq.log 2>&1; echo "rc=$?"
tail -12 gpurun_out/r02ar_pytest_vq.log
