#!/bin/bash
# solve / factor / whole-step time against the leaf size of the nested-dissection tree (1M-triangle vector problem)
cd "$(dirname "$0")/.."
for L in 16 24 32 48 64 96; do
  echo "== leaf_elems $L"
  WS_LEAF_ELEMS=$L python tools/solve_only.py --solves 20 --factors 2 --leaf $L 2>&1 | tail -1
  WS_LEAF_ELEMS=$L python bench.py --lite --steps 5 --warmup 2 --no-clock-sampler 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', round(d['ms_per_step'],1), 'ops', d['eig']['op_applications'])"
done
