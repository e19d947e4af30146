#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_measure.py -m gpu -x -q -k "fused or config4" 2>&1 | tail -3
for lib in ${LIBS:-libmolar_b200.so}; do
for o in ${FIT_OPTS:-"fused_fit=4"}; do
  echo "== $lib $o"
  MOLAR_B200_PLUGIN=$PWD/molar_b200/lib/$lib timeout 120 python bench.py --workload fit500k --steps 5 --warmup 3 --no-cpu --no-e2e --opts "$o" 2>&1 | tail -1 | cut -c1-200
done; done
