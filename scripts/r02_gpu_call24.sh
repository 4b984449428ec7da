#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "do_not_depend_on_the_window or (window_forward and sw)" > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02w_pytest.log
out=gpurun_out/r02w_bench_ab.txt
: > $out
for cfg in "SJD_ATTN_SW_AUTO=0" "SJD_ATTN_SW_AUTO=1" "SJD_ATTN_SW_AUTO=0" "SJD_ATTN_SW_AUTO=1"; do
  echo "== $cfg" >> $out
  env $cfg $T 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference 2>/dev/null | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'ms_per_nfe', d['ms_per_nfe'], 'nfe', d['nfe_per_image'], 'chain frac', d['roofline']['frac'])
" >> $out
done
cat $out
