#!/bin/bash
# 8-GPU lines of the three multi-GPU workloads (one rank per GPU, NCCL scalar gather inside the library)
N=${1:-8}; TAG=${2:-r4}
mkdir -p gpurun_out
for w in search1m fit500k pipeline1m; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload $w --no-cpu > gpurun_out/bench_${TAG}_${N}gpu_${w}.json 2> gpurun_out/bench_${TAG}_${N}gpu_${w}.err
  cut -c1-260 gpurun_out/bench_${TAG}_${N}gpu_${w}.json; tail -2 gpurun_out/bench_${TAG}_${N}gpu_${w}.err | cut -c1-200
done
