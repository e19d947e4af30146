#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2h}
timeout 400 python -m pytest tests/test_gpu_measure.py tests/test_gpu_search.py -m gpu -x -q -k "multi_chunk or multipass" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_search1m.json 2> gpurun_out/bench_${TAG}_search1m.err
cut -c1-400 gpurun_out/bench_${TAG}_search1m.json; tail -2 gpurun_out/bench_${TAG}_search1m.err
timeout 500 python tools/bench_extra.py > gpurun_out/bench_extra_${TAG}.jsonl 2> gpurun_out/bench_extra_${TAG}.err; tail -2 gpurun_out/bench_extra_${TAG}.err; wc -l gpurun_out/bench_extra_${TAG}.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
