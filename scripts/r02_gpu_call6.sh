#!/bin/bash
# round 2, GPU call 6: validate the verify-kernel changes, re-time, capture launch list
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "gpu suites rc=$?"
tail -6 gpurun_out/r02f_pytest_gpu.log | cut -c1-300
$T 120 python scripts/profile_trips.py --time --trips 20 > gpurun_out/r02f_trip_times.txt 2>&1; grep "W=" gpurun_out/r02f_trip_times.txt
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02f_smoke.log
$T 800 python bench.py --steps 3 --warmup 3 --cpu-budget 8 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
cat gpurun_out/r02f_bench.json | cut -c1-900; tail -2 gpurun_out/r02f_bench.err
$T 300 python bench.py --impl reference --steps 1 --warmup 0 --cpu-budget 8 > gpurun_out/r02f_bench_reference.json 2>/dev/null; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/r02f_bench_reference.json
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_chain|attn_|embed_rmsnorm|verify_|gather_rows" --launch-skip 10 -c 300 --csv --log-file gpurun_out/r02f_launches.csv python scripts/profile_trips.py --trips 5 > gpurun_out/r02f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$T 400 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/r02f_full_verify -f python scripts/profile_trips.py --trips 2 > gpurun_out/r02f_ncu_verify.log 2>&1; echo "ncu verify rc=$?"
