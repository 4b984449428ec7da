#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 1500 python -m pytest tests -q -m gpu > gpurun_out/r02bb_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
tail -4 gpurun_out/r02bb_pytest_gpu.log
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bb_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02bb_smoke.log
out=gpurun_out/r02bb_bench_ab.txt
: > $out
for cfg in "SJD_ATTN_SW_AUTO=0 SJD_ZERO_COPY=0" "SJD_ATTN_SW_CLUSTER=0" "SJD_ATTN_SW_CLUSTER=4"; do
  echo "== $cfg" >> $out
  env $cfg $T 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms_per_nfe', d['ms_per_nfe'], 'nfe', d['nfe_per_image'], 'chain frac', d['roofline']['frac'])
" >> $out
done
cat $out
