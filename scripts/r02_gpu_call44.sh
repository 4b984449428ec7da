#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/trip_gap.py > gpurun_out/r02aq_trip_gap.txt 2>&1; echo rc=$?
tail -3 gpurun_out/r02aq_trip_gap.txt
