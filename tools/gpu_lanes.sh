#!/bin/bash
# lane kernel bring-up: parity suite, then bench with the lane kernel and with the mask kernel
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -15
for o in "" "lane_kernel=0" ; do
  timeout 300 python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e --opts "$o" 2>gpurun_out/bench_${TAG}.err | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-40s kernel %.3f ms  total %.3f ms/frame' % (d.get('opts'), d.get('search_kernel_ms',0), d.get('ms_per_frame',0)))"
done
tail -3 gpurun_out/bench_${TAG}.err
