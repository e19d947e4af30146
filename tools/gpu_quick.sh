#!/bin/bash
# quick GPU iteration: parity tests (optional filter), bench line, ncu metrics of the search kernel
TAG=${1:-x}
KEXPR=${2:-"cells or filter or config or partial or degenerate or fixture or ids"}
mkdir -p gpurun_out
if [ "$KEXPR" != "none" ]; then
python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "$KEXPR" 2>&1 | tail -8
fi
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
cut -c1-900 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
