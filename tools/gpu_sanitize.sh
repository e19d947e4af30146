#!/bin/bash
# compute-sanitizer memcheck (+ racecheck with "race" as first argument) over the small-config smoke of every kernel family
mkdir -p gpurun_out
TAG=${2:-r3}
if [ "$1" == "race" ]; then
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_racecheck_${TAG}.txt 2>&1; echo "rc=$?" >> gpurun_out/sanitize_racecheck_${TAG}.txt
tail -4 gpurun_out/sanitize_racecheck_${TAG}.txt
else
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck_${TAG}.txt 2>&1; echo "rc=$?" >> gpurun_out/sanitize_memcheck_${TAG}.txt
tail -4 gpurun_out/sanitize_memcheck_${TAG}.txt
fi
