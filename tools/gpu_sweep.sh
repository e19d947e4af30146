#!/bin/bash
# sweep traversal-grid options on the headline workload (tuning only)
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${1:-x}.jsonl
: > $OUT
for o in "" "subdiv_x=2,subdiv_y=2,subdiv_z=3,slice_x=4" "subdiv_x=2,subdiv_y=3,subdiv_z=2,slice_x=4" "subdiv_x=3,subdiv_y=2,subdiv_z=2,slice_x=3" "subdiv_x=2,subdiv_y=3,subdiv_z=3,slice_x=4" "subdiv_x=3,subdiv_y=3,subdiv_z=3,slice_x=4" "subdiv_x=3,subdiv_y=3,subdiv_z=3,slice_x=2" "subdiv_x=3,subdiv_y=3,subdiv_z=4,slice_x=3" "subdiv_x=2,subdiv_y=4,subdiv_z=4,slice_x=4" "subdiv_x=3,subdiv_y=2,subdiv_z=3,slice_x=3"; do
  python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e --opts "$o" 2>&1 | tail -1 >> $OUT
done
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-50s kernel %.3f ms  total %.3f ms/frame' % (d['opts'], d['search_kernel_ms'], d['ms_per_frame']))"
