#!/bin/bash
# ncu --set full capture of the lane kernel on the headline workload (+ optional options)
TAG=${1:-x}; OPTS=${2:-}
mkdir -p gpurun_out
MB_DEBUG_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e --opts "$OPTS" 2>gpurun_out/bench_${TAG}.err | tail -1 | cut -c1-300
grep -m1 "batch_search" gpurun_out/bench_${TAG}.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_ -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e --opts "$OPTS" > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
