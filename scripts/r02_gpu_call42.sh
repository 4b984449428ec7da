#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
for cfg in 1 4 5; do
$T 600 python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02ao_bench_config$cfg.json 2> gpurun_out/r02ao_bench_config$cfg.err; echo "bench config $cfg rc=$?"
done
python - <<'PY'
import json
for c in (1, 4, 5):
    d = json.loads([l for l in open(f'gpurun_out/r02ao_bench_config{c}.json') if l.startswith('{')][0])
    print(c, d['value'], d['ms_per_nfe'], [(r['window'], r['tokens_per_s'], r['ms_per_nfe']) for r in d.get('window_sweep', [])])
PY
$T 400 ncu --set full --clock-control none --import-source on -k regex:"attn_sw_kernel" --launch-skip 40 --launch-count 2 -o gpurun_out/r02ao_full_attn -f python scripts/profile_trips.py --trips 2 > gpurun_out/r02ao_ncu_full.log 2>&1; echo "ncu full rc=$?"
