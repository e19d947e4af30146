#!/bin/bash
mkdir -p gpurun_out
TAG=$1; OPTS_A=$2; OPTS_B=$3
ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e --opts "$OPTS_A" > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --frames 4 --no-cpu --no-e2e --opts "$OPTS_B" > gpurun_out/launches_${TAG}.log 2>&1
tail -2 gpurun_out/launches_${TAG}.log
