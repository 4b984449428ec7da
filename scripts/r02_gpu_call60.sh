#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 200 python scripts/trip_gap.py 2>&1 | tail -1 > gpurun_out/r02bg_trip_gap.txt; cat gpurun_out/r02bg_trip_gap.txt
$T 900 python -m pytest tests -q -m gpu > gpurun_out/r02bg_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -3 gpurun_out/r02bg_pytest_gpu.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bg_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02bg_smoke.log
