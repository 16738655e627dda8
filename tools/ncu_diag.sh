#!/bin/bash
# diagnostic: does ncu work at all on this box, and on our library (short commands, hard timeouts)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== step 1: torch only"; date +%s
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 \
   python -c "import torch; x=torch.ones(1000,device='cuda'); print((x*2).sum().item())" > gpurun_out/diag1.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/diag1.log; date +%s
echo "== step 2: solve_only 20k elements, no csv"
python tools/solve_only.py --elements 20000 --solves 1 > gpurun_out/diag2_plain.log 2>&1 &&
timeout -s KILL 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 \
   python tools/solve_only.py --elements 20000 --solves 1 > gpurun_out/diag2.log 2>&1
echo "rc=$?"; grep -c "PROF== Profiling" gpurun_out/diag2.log; tail -3 gpurun_out/diag2.log; date +%s
