#!/bin/bash
# compute-sanitizer passes over one small train + eval step: racecheck (shared-memory hazards of the
# mbarrier / bulk-copy / TMEM protocols), synccheck (barrier misuse) and memcheck.  Run under gpurun; logs -> gpurun_out/.
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool" > gpurun_out/r2_sanitizer_$tool.log
  ( time timeout 420 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_step.py 6 700 ) >> gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/r2_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_step ok|exit code" gpurun_out/r2_sanitizer_$tool.log | head -5
done
