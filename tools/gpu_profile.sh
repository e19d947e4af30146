#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, one full capture of the search kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r1}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
tail -5 gpurun_out/bench_${TAG}.err
# launch list (per-launch device time, cold cache, serialised): same command, short
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-cpu \
    > gpurun_out/launches_${TAG}.log 2>&1
tail -3 gpurun_out/launches_${TAG}.log
# full capture of the dominant kernel (3 launches after warm-up)
ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 2 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -3 gpurun_out/prof_${TAG}.log
ls -la gpurun_out
