#!/bin/bash
# Runs on the GPU box (under gpurun): smoke, bench line, ncu launch list, one full capture of the search kernel.
mkdir -p gpurun_out
TAG=${1:-r1}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cut -c1-1600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_${TAG}.err
cut -c1-400 gpurun_out/bench_${TAG}_reference.json
# launch list (per-launch device time, cold cache, serialised): same command, shorter
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-cpu --no-e2e \
    > gpurun_out/launches_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.log
ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
