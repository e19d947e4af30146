#!/bin/bash
# quick GPU iteration: subset of parity tests, bench line, ncu metrics of the search kernel
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "cells or filter or config2 or partial or degenerate" 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cut -c1-900 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
