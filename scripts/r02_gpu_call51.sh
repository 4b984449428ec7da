#!/bin/bash
# final record of the round: headline bench line (defaults) + the reference arm
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python bench.py > gpurun_out/r02ax_bench.json 2> gpurun_out/r02ax_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/r02ax_bench.json | cut -c1-200
$T 300 python bench.py --impl reference --steps 1 --warmup 0 --cpu-budget 8 > gpurun_out/r02ax_bench_reference.json 2>/dev/null; echo "ref arm rc=$?"; cut -c1-300 gpurun_out/r02ax_bench_reference.json
