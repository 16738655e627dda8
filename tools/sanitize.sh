#!/bin/bash
# compute-sanitizer record of one small pass over the hot path (memcheck, racecheck, synccheck) -> gpurun_out/sanitize_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export WS_NO_GRAPH=1            # kernels launched directly (the sanitizer instruments launches, not graph replays)
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|scalar|vector|sweep|interior" gpurun_out/sanitize_$tool.log | tail -12
done
