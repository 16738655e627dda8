#!/bin/bash
# end-of-round check on the GPU box: full parity suite twice (flag-synchronised sweeps must be stable),
# smoke, the per-launch solve list with DRAM bytes, the bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
export WS_NO_GRAPH=1
KERN="python tools/solve_only.py --solves 1"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k 'regex:k_fwd|k_bwd|k_spmv' --csv --log-file gpurun_out/solve_list.csv $KERN > gpurun_out/ncu_solve.log 2>&1
echo "solve list rc=$?"
unset WS_NO_GRAPH
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1
tail -c 400 gpurun_out/bench_final.log
