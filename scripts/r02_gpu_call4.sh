#!/bin/bash
# round 2, GPU call 4: suites with device Philox + Anole modes, bench lines for every config, ncu evidence
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "gpu suites rc=$?"
tail -12 gpurun_out/r02d_pytest_gpu.log | cut -c1-300
$T 120 python scripts/profile_trips.py --time --trips 10 > gpurun_out/r02d_trip_times.txt 2>&1
$T 120 python scripts/profile_trips.py --time --trips 10 --host-noise >> gpurun_out/r02d_trip_times.txt 2>&1
cat gpurun_out/r02d_trip_times.txt | grep "W="
$T 800 python bench.py --steps 3 --warmup 3 --cpu-budget 8 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
cat gpurun_out/r02d_bench.json | cut -c1-700; tail -2 gpurun_out/r02d_bench.err
for c in 1 5 4; do
  $T 600 python bench.py --config $c --steps 1 --warmup 1 > gpurun_out/r02d_bench_config$c.json 2> gpurun_out/r02d_bench_config$c.err; echo "config $c rc=$?"
  cat gpurun_out/r02d_bench_config$c.json | cut -c1-600; tail -2 gpurun_out/r02d_bench_config$c.err
done
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02d_launches.csv python scripts/profile_trips.py --trips 3 > gpurun_out/r02d_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$T 500 ncu --set full --clock-control none --import-source on -k regex:"gemm_chain|attn_window" --launch-skip 70 --launch-count 6 -o gpurun_out/r02d_full -f python scripts/profile_trips.py --trips 2 > gpurun_out/r02d_ncu_full.log 2>&1; echo "ncu full rc=$?"
$T 400 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/r02d_full_verify -f python scripts/profile_trips.py --trips 2 > gpurun_out/r02d_ncu_verify.log 2>&1; echo "ncu verify rc=$?"
ls -la gpurun_out | grep r02d
