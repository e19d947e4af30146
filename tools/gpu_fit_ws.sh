#!/bin/bash
# warp-specialised Kabsch kernel: parity, then throughput against the default two-kernel path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_measure.py -m gpu -x -q -k "fused or config4" 2>&1 | tail -4
for o in ${FIT_OPTS:-"fused_fit=0" "fused_fit=4" "fused_fit=4,fit_lag=12" "fused_fit=4,fit_lag=20" "fused_fit=4,fit_lag=32"}; do
  echo "== $o"
  timeout 120 python bench.py --workload fit500k --steps 5 --warmup 3 --no-cpu --no-e2e --opts "$o" 2>/dev/null | cut -c1-200
done
