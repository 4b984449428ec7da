#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02bh_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02bh_smoke.log
SJD_NVTX=1 $T 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden_loops or real_stack or anole_adaptor or emu3_adaptor" > gpurun_out/r02bh_pytest.log 2>&1; echo "pytest (NVTX on) rc=$?"; tail -2 gpurun_out/r02bh_pytest.log
