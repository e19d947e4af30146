#!/bin/bash
# one iteration on the search kernel: a parity subset, then the tuning bench line(s)
TAG=${1:-x}; KEXPR=${2:-"orthorhombic or triclinic_config3 or partial or multipass or config2 or double_and_within"}; shift; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py -m gpu -x -q -k "$KEXPR" --durations=5 2>&1 | tail -12
for o in "$@"; do
  MB_DEBUG_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e --opts "$o" 2>gpurun_out/bench_${TAG}.err | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-40s kernel %.3f ms  total %.3f ms/frame' % (d.get('opts'), d.get('search_kernel_ms',0), d.get('ms_per_frame',0)))"
  grep -m1 batch_search gpurun_out/bench_${TAG}.err | cut -c1-200
done
