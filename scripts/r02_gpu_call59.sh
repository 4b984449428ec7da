#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python scripts/host_profile.py > gpurun_out/r02bf_host_profile.txt 2>&1; echo rc=$?
head -45 gpurun_out/r02bf_host_profile.txt
