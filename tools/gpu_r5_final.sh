#!/bin/bash
# round-2 final evidence run (one GPU): remaining parity tests, smoke, bench lines of every workload + reference arm,
# next-row measurements, launch list, ncu captures of the pair kernel, the neighbour-list modes, the PBC reductions
# and the warp-specialised Kabsch kernel
mkdir -p gpurun_out
TAG=${1:-r5}
timeout 900 python -m pytest tests/test_gpu_measure.py tests/test_gpu_measure_pbc.py tests/test_gpu_connect.py tests/test_gpu_traj.py -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_search1m.json 2> gpurun_out/bench_${TAG}_search1m.err
cut -c1-330 gpurun_out/bench_${TAG}_search1m.json; tail -1 gpurun_out/bench_${TAG}_search1m.err
for w in search100k fit500k pipeline1m; do
  timeout 400 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
  cut -c1-330 gpurun_out/bench_${TAG}_$w.json; tail -1 gpurun_out/bench_${TAG}_$w.err
done
timeout 200 python bench.py --workload fit500k --steps 10 --warmup 3 --no-cpu --no-e2e --opts fused_fit=4 2>/dev/null | cut -c1-200 | tee gpurun_out/bench_${TAG}_fit500k_ws.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_search1m_reference.json 2> gpurun_out/bench_${TAG}_reference.err
cut -c1-300 gpurun_out/bench_${TAG}_search1m_reference.json
timeout 500 python tools/bench_extra.py > gpurun_out/bench_extra_${TAG}.jsonl 2> gpurun_out/bench_extra_${TAG}.err; tail -2 gpurun_out/bench_extra_${TAG}.err; wc -l gpurun_out/bench_extra_${TAG}.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --frames 4 --no-cpu --no-e2e \
    > gpurun_out/launches_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_cells_kernel -s 4 -c 1 \
    -o gpurun_out/prof_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e \
    > gpurun_out/prof_${TAG}.log 2>&1
tail -1 gpurun_out/prof_${TAG}.log
timeout 400 ncu --set full --clock-control none -k regex:'search_cells_kernel|csr_count_kernel|csr_fill_kernel' -c 6 \
    -o gpurun_out/prof_nl_${TAG} -f python tools/bench_extra.py --only connect > gpurun_out/prof_nl_${TAG}.log 2>&1
tail -1 gpurun_out/prof_nl_${TAG}.log
timeout 400 ncu --set full --clock-control none -k regex:'center_pbc_kernel|tensor_kernel' -c 4 \
    -o gpurun_out/prof_pbc_${TAG} -f python tools/bench_extra.py --only pbc --reps 2 > gpurun_out/prof_pbc_${TAG}.log 2>&1
tail -1 gpurun_out/prof_pbc_${TAG}.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fit_ws_kernel -s 3 -c 1 \
    -o gpurun_out/prof_fit_ws_${TAG} -f python bench.py --workload fit500k --steps 1 --warmup 3 --frames 64 --no-cpu --no-e2e --opts fused_fit=4 \
    > gpurun_out/prof_fit_ws_${TAG}.log 2>&1
tail -1 gpurun_out/prof_fit_ws_${TAG}.log
