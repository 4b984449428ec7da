#!/bin/bash
# round 2, GPU call 5: 3-stage attention A/B, Anole residual sets, sanitizer, launch list
mkdir -p gpurun_out
T="timeout -k 10"
out=gpurun_out/r02e_attn_stages.txt
: > $out
for cfg in "SJD_ATTN_STAGES=2" "SJD_ATTN_STAGES=3"; do
  echo "== $cfg" >> $out
  env $cfg $T 200 python scripts/chain_time.py 8 8,16,32 2>&1 | grep "W=" >> $out
done
cat $out
SJD_ATTN_STAGES=3 $T 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -k "window_forward or full_width or real_stack or end_to_end or config2" > gpurun_out/r02e_pytest_stages3.log 2>&1; echo "stages3 forward tests rc=$?"
tail -4 gpurun_out/r02e_pytest_stages3.log | cut -c1-300
$T 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "gpu suites rc=$?"
tail -6 gpurun_out/r02e_pytest_gpu.log | cut -c1-300
bash scripts/sanitize.sh memcheck
bash scripts/sanitize.sh racecheck
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sjd::" --launch-skip 140 -c 300 --csv --log-file gpurun_out/r02e_launches.csv python scripts/profile_trips.py --trips 4 > gpurun_out/r02e_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
