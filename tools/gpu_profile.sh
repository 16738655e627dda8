#!/bin/bash
# Runs on the GPU box (gpurun): plain bench, ncu launch list of the same command, one
# `ncu --set full` capture of the hot kernels (assembly, SpMV, solve sweeps).  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
KERN="python tools/solve_only.py --solves 1"
$BENCH > gpurun_out/plain_bench_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
$KERN > gpurun_out/plain_kern_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k 'regex:k_assemble_rows|k_spmv_stream|k_fwd_gemv|k_bwd_gemvT|k_fwd_small|k_bwd_small' -c 120 \
    -f -o gpurun_out/prof_$TAG $KERN > gpurun_out/ncu_kern_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out
