#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_racecheck_r2.txt 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_r2.txt
tail -4 gpurun_out/sanitize_racecheck_r2.txt
