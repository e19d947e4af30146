#!/bin/bash
# Kabsch batch tuning lines only (no tests)
for o in "$@"; do
  timeout 300 python bench.py --workload fit500k --steps 5 --warmup 3 --no-cpu --no-e2e --opts "$o" 2>gpurun_out/fit_err.txt | tail -1 | cut -c1-130
done
