#!/bin/bash
# Runs on the GPU box (gpurun).  1) `ncu --set full` capture of the hot kernels (assembly, SpMV,
# solve sweeps, fused small-front factorisation) on the 1M-element vector problem, reduced ON THE
# BOX to csv pages (the .ncu-rep itself is too large to travel back); 2) per-launch list of one
# shift-invert solve with DRAM bytes; 3) ncu launch list of one bench.py step.
# Every profiler step sits behind the plain run of the same command and a hard timeout.
# (WS_NO_GRAPH=1: same kernels, launched directly instead of through CUDA graphs -- ncu on the
#  graph-replayed bench did not make progress in 25 minutes.)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01}
export WS_NO_GRAPH=1
KERN="python tools/solve_only.py --solves 1"
REP=/tmp/prof_$TAG
$KERN > gpurun_out/plain_kern_$TAG.log 2>&1 &&
timeout -s KILL 700 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_assemble_rows|k_spmv_stream|k_fwd_front|k_bwd_front|k_fwd_tile|k_bwd_tile|k_front_fused' -c 48 \
    -f -o $REP $KERN > gpurun_out/ncu_kern_$TAG.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_kern_$TAG.log
if [ -f $REP.ncu-rep ]; then
  ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
  ncu -i $REP.ncu-rep --page details --csv > gpurun_out/prof_${TAG}_details.csv 2>/dev/null
  ls -la $REP.ncu-rep
fi
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k 'regex:k_fwd|k_bwd|k_spmv' --csv --log-file gpurun_out/solve_list.csv $KERN > gpurun_out/ncu_solve.log 2>&1
echo "solve list rc=$?"
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-clock-sampler --lite"
$BENCH > gpurun_out/plain_bench_$TAG.log 2>&1 &&
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
# 4) wide Krylov basis (Nmax = 80): the restart rotation as a DMMA GEMM (k_rotate_mma) and the CGS kernels -- the fp64 / tensor
#    pipe figures north_star asks for
WIDE="python tools/wide_basis_solve.py"
$WIDE > gpurun_out/plain_wide_$TAG.log 2>&1 &&
timeout -s KILL 600 ncu --set full --clock-control none -k 'regex:k_rotate_mma|k_dots_partial|k_update' -c 12 --launch-skip 600 \
    -f -o ${REP}_wide $WIDE > gpurun_out/ncu_wide_$TAG.log 2>&1
echo "wide capture rc=$?"; cat gpurun_out/plain_wide_$TAG.log
if [ -f ${REP}_wide.ncu-rep ]; then
  ncu -i ${REP}_wide.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_wide_raw.csv 2>/dev/null
fi
du -sh gpurun_out
