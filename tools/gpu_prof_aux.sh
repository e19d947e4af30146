#!/bin/bash
# ncu --set full of the kernels around the pair kernel (binning, scatter, scans) and of the Kabsch two-kernel path
TAG=${1:-r4}
timeout 600 ncu --set full --clock-control none -k regex:'bin_atoms_kernel|scatter_kernel|scan_apply_kernel|merge_pair_tail_kernel' -s 16 -c 4 \
    -o gpurun_out/prof_aux_search_${TAG} -f python bench.py --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e > gpurun_out/prof_aux_${TAG}.log 2>&1
tail -1 gpurun_out/prof_aux_${TAG}.log
timeout 600 ncu --set full --clock-control none -k regex:'fit_moments_kernel|superpose_rmsd_kernel' -s 4 -c 2 \
    -o gpurun_out/prof_aux_fit_${TAG} -f python bench.py --workload fit500k --steps 1 --warmup 3 --frames 64 --no-cpu --no-e2e > gpurun_out/prof_aux_fit_${TAG}.log 2>&1
tail -1 gpurun_out/prof_aux_fit_${TAG}.log
timeout 600 ncu --set full --clock-control none -k regex:'moments1_kernel|search_cells_kernel' -s 6 -c 2 \
    -o gpurun_out/prof_aux_pipe_${TAG} -f python bench.py --workload pipeline1m --steps 1 --warmup 3 --frames 2 --no-cpu --no-e2e > gpurun_out/prof_aux_pipe_${TAG}.log 2>&1
tail -1 gpurun_out/prof_aux_pipe_${TAG}.log
# launch list of the default bench command (per-launch times are cold-cache and serialised: the SHARE must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --frames 4 --no-cpu --no-e2e > gpurun_out/launches_${TAG}.log 2>&1
tail -1 gpurun_out/launches_${TAG}.log
timeout 600 ncu --set full --clock-control none -k regex:'center_pbc_kernel|tensor_kernel|reduce_many' -c 6 \
    -o gpurun_out/prof_aux_pbc_${TAG} -f python tools/bench_extra.py --only pbc > gpurun_out/prof_aux_pbc_${TAG}.log 2>&1
tail -1 gpurun_out/prof_aux_pbc_${TAG}.log
