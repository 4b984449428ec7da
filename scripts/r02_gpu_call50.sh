#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02aw_zero_copy.txt
: > $out
for cfg in "SJD_ZERO_COPY=0" "SJD_ZERO_COPY=1" "SJD_ZERO_COPY=0" "SJD_ZERO_COPY=1"; do
  echo "== $cfg" >> $out
  env $cfg $T 200 python scripts/trip_gap.py 2>&1 | tail -1 >> $out
done
cat $out
$T 1500 python -m pytest tests -q -m gpu > gpurun_out/r02aw_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02aw_pytest_gpu.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02aw_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02aw_smoke.log
