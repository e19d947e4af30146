#!/bin/bash
# usage: gpurun_retry.sh OUTFILE TIMEOUT 'command'   — retries while the pod has no free slot (nothing is charged then)
OUT=$1; TMO=$2; CMD=$3; GP=${4:-1}
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $GP --timeout $TMO -- "$CMD" > $OUT 2>&1
  if grep -q "status=transient\|retry in a few minutes\|no box" $OUT; then sleep 90; continue; fi
  break
done
