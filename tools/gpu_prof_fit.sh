#!/bin/bash
# ncu --set full of the Kabsch batch kernels: usage gpu_prof_fit.sh TAG KERNEL_REGEX OPTS [FRAMES]
TAG=$1; KRE=$2; OPTS=$3; FR=${4:-32}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 \
    -o gpurun_out/prof_fit_${TAG} -f python bench.py --workload fit500k --steps 1 --warmup 3 --frames $FR --no-cpu --no-e2e --opts "$OPTS" \
    > gpurun_out/prof_fit_${TAG}.log 2>&1
tail -2 gpurun_out/prof_fit_${TAG}.log
