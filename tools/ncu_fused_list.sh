#!/bin/bash
# per-launch device time of the fused small-front kernel (one factorisation, 1M-element vector problem)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/solve_only.py --solves 1 --factors 1"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none \
   -k 'regex:k_front_fused' --csv --log-file gpurun_out/fused_list.csv $CMD > gpurun_out/ncu_fused.log 2>&1
echo rc=$?
grep -v "^==" gpurun_out/fused_list.csv | python -c "
import csv,sys
from collections import OrderedDict
L=OrderedDict()
for r in csv.DictReader(sys.stdin):
    d=L.setdefault(r['ID'],{'grid':r['Grid Size'],'block':r['Block Size']})
    d[r['Metric Name']]=r['Metric Value']+' '+r['Metric Unit']
for k,d in L.items(): print(d)
"
