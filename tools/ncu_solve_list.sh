#!/bin/bash
# launch list (device time per kernel) of one factor + 2 solves on the 1M-element vector problem
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export WS_NO_GRAPH=1
CMD="python tools/solve_only.py --solves 1"
$CMD > gpurun_out/plain_solve.log 2>&1 &&
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k 'regex:k_fwd|k_bwd|k_spmv' --csv --log-file gpurun_out/solve_list.csv $CMD > gpurun_out/ncu_solve.log 2>&1
echo rc=$?; cat gpurun_out/plain_solve.log
