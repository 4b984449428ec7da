#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -q -m gpu > gpurun_out/r02be_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02be_pytest_gpu.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02be_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02be_smoke.log
SJD_ATTN=sw $T 150 python scripts/chain_time.py 8 32 2>&1 | grep "W="
