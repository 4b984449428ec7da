#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 bash scripts/sanitize.sh memcheck > gpurun_out/r02af_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/r02af_sanitize_memcheck.log
timeout -k 10 900 bash scripts/sanitize.sh racecheck > gpurun_out/r02af_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -8 gpurun_out/r02af_sanitize_racecheck.log
