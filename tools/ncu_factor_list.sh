#!/bin/bash
# per-launch device time of one numeric factorisation on the 1M-element vector problem
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/solve_only.py --solves 1 --factors 1"
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   -k 'regex:k_front_fused|k_gemm_tiles|k_panel|k_extend_add|k_diag_inverse|k_scatter|k_recip' --csv --log-file gpurun_out/factor_list.csv $CMD > gpurun_out/ncu_factor.log 2>&1
echo rc=$?
python tools/summarize_launches.py gpurun_out/factor_list.csv 20
python - <<'PY'
import csv
from collections import OrderedDict
lines=[l for l in open('gpurun_out/factor_list.csv') if not l.startswith('==')]
L=OrderedDict()
for r in csv.DictReader(lines):
    d=L.setdefault(r['ID'],{'name':r['Kernel Name'].split('(')[0][-28:],'grid':r['Grid Size'],'block':r['Block Size']})
    d[r['Metric Name']]=(float(r['Metric Value'].replace(',','')),r['Metric Unit'])
for k,d in L.items():
    if 'fused' in d['name']:
        t=d['gpu__time_duration.sum']; tu=t[0]/1000 if t[1] in('ns','nsecond') else t[0]
        print(d['name'],d['grid'],d['block'],'%.1f us'%tu, d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'), d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed'))
PY
