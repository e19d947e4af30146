#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck_r2.txt 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_r2.txt
tail -4 gpurun_out/sanitize_memcheck_r2.txt
