#!/bin/bash
# round 2, GPU call 1: new parity tests first, then the whole GPU suite, chain experiments, stamps, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r02_smi.txt 2>&1
python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu > gpurun_out/r02_pytest_new.log 2>&1; echo "new tests rc=$?" 
tail -5 gpurun_out/r02_pytest_new.log
python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log
bash scripts/r02_chain_experiments.sh > /dev/null 2>&1
cat gpurun_out/r02_chain_experiments.txt
python scripts/gemm_stamps.py 32 > gpurun_out/r02_gemm_stamps_w32.txt 2>&1; head -6 gpurun_out/r02_gemm_stamps_w32.txt
timeout 900 python bench.py --steps 2 --warmup 3 --cpu-budget 10 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
cat gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
