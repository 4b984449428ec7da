#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
for sc in 0.05 0.2; do
  $T 300 python bench.py --steps 2 --warmup 1 --logit-scale $sc --top-k 8192 > gpurun_out/r02l_bench_demo_scale$sc.json 2> gpurun_out/r02l_demo.err; echo "demo scale $sc rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02l_bench_demo_scale$sc.json'))
print({k:d[k] for k in ('value','accepted_tokens_per_iter','nfe_per_image','ms_per_nfe')}, d['e2e']['value'])
PY
done
tail -3 gpurun_out/r02l_demo.err
