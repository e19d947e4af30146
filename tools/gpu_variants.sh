#!/bin/bash
# bench line of the default library and of each tuning build under molar_b200/lib/variants (MOLAR_B200_PLUGIN)
for v in "" "$@"; do
  if [ -n "$v" ]; then export MOLAR_B200_PLUGIN=$PWD/molar_b200/lib/variants/lib_$v.so; fi
  echo -n "variant=$v  "
  MB_DEBUG_TIMING=0 timeout 300 python bench.py --steps 3 --warmup 3 --frames 8 --no-cpu --no-e2e 2>/dev/null | tail -1 | cut -c1-160
done
