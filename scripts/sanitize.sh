#!/bin/bash
# compute-sanitizer over the smoke decode (GPU box):  bash scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
# The kernels lean on cross-CTA flags, __nanosleep polls and counters that re-arm themselves (gemm_chain_kernel's fin /
# tile_arrive words, verify_kernel's sync_ws): memcheck catches out-of-bounds / misaligned accesses of the parked partials
# and the KV cache, racecheck the shared-memory hand-offs (epilogue staging tile, softmax exchange), synccheck the named
# barriers.  Output: gpurun_out/r02_sanitize_<tool>.log (a summary line is printed).
tool=${1:-memcheck}
mkdir -p gpurun_out
log=gpurun_out/r02_sanitize_$tool.log
timeout -k 10 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 3 \
  python -c "import __graft_entry__ as g; g.smoke()" > $log 2>&1
rc=$?
echo "compute-sanitizer $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke:' $log | tr '\n' ' ')"
exit $rc
