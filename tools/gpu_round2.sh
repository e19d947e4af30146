#!/bin/bash
# end-of-session evidence run: bench lines for every workload, reference arm, ncu launch list + one full capture
mkdir -p gpurun_out
TAG=${1:-r2}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in search100k fit500k pipeline1m; do
  timeout 400 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
  cut -c1-330 gpurun_out/bench_${TAG}_$w.json; tail -1 gpurun_out/bench_${TAG}_$w.err
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
cut -c1-300 gpurun_out/bench_${TAG}_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-cpu --no-e2e \
    > gpurun_out/launches_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
timeout 300 ncu --set full --clock-control none -k regex:"dcd_unpack|xtc_|center_pbc|tensor_kernel" -c 12 \
    -o gpurun_out/prof_extra_${TAG} -f python tools/bench_extra.py --atoms 1000000 --frames 8 --reps 2 \
    > gpurun_out/prof_extra_${TAG}.log 2>&1
tail -1 gpurun_out/prof_extra_${TAG}.log
