#!/bin/bash
# round-3 iteration on the search kernel: parity (search tests), tuning bench lines, optional ncu capture
TAG=${1:-x}; KEXPR=${2:-"not 1m"}; PROF=${3:-0}; shift; shift; shift
mkdir -p gpurun_out
if [ "$KEXPR" != "none" ]; then
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "$KEXPR" --durations=3 2>&1 | tail -8
fi
for o in "$@"; do
  oo=$o; [ "$o" == "default" ] && oo=""
  MB_DEBUG_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e --opts "$oo" 2>gpurun_out/bench_${TAG}.err | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-40s kernel %.3f ms  total %.3f ms/frame' % (d.get('opts'), d.get('search_kernel_ms',0), d.get('ms_per_frame',0)))"
done
if [ "$PROF" != "0" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
fi
