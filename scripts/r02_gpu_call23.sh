#!/bin/bash
# full GPU suite with the new default attention + headline bench
mkdir -p gpurun_out
T="timeout -k 10"
$T 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -5 gpurun_out/r02v_pytest_gpu.log
$T 900 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/r02v_bench.json | cut -c1-1500
