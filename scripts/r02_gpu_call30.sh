#!/bin/bash
# evidence for the final round-2 kernels: trip times, smoke, launch list, --set full captures, bench lines of every config
mkdir -p gpurun_out
T="timeout -k 10"
$T 120 python scripts/profile_trips.py --time --trips 20 > gpurun_out/r02ac_trip_times.txt 2>&1; grep "W=" gpurun_out/r02ac_trip_times.txt
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02ac_smoke.log
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_chain|attn_|embed_rmsnorm|verify_|gather_rows" --launch-skip 10 -c 300 --csv --log-file gpurun_out/r02ac_launches.csv python scripts/profile_trips.py --trips 5 > gpurun_out/r02ac_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$T 500 ncu --set full --clock-control none --import-source on -k regex:"attn_sw_kernel|gemm_chain" --launch-skip 70 --launch-count 4 -o gpurun_out/r02ac_full -f python scripts/profile_trips.py --trips 2 > gpurun_out/r02ac_ncu_full.log 2>&1; echo "ncu full rc=$?"
for cfg in 1 4 5; do
$T 600 python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference > gpurun_out/r02ac_bench_config$cfg.json 2> gpurun_out/r02ac_bench_config$cfg.err; echo "bench config $cfg rc=$?"
grep '^{' gpurun_out/r02ac_bench_config$cfg.json | cut -c1-400
done
ls -la gpurun_out/r02ac*
