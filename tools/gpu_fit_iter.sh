#!/bin/bash
# Kabsch batch iteration: parity of the fit tests, then the config-4 bench line for each option string
timeout 600 python -m pytest tests/test_gpu_measure.py -m gpu -x -q -k "fit" --durations=3 2>&1 | tail -6
for o in "$@"; do
  timeout 300 python bench.py --workload fit500k --steps 5 --warmup 3 --no-cpu --no-e2e --opts "$o" 2>gpurun_out/fit_err.txt | tail -1 | cut -c1-160
done
