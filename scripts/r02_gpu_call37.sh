#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 1500 python -m pytest tests -q -m gpu > gpurun_out/r02aj_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02aj_pytest_gpu.log
