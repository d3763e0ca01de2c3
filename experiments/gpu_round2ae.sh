#!/bin/bash
# round 2, call AE: compute-sanitizer (memcheck, racecheck, synccheck) on smoke() + a multi-candidate CCX call
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 140 compute-sanitizer --tool $tool --print-limit 20 python experiments/sanitize_ccx.py > gpurun_out/r2ae_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2ae_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=|sanitize target ok|smoke ok" gpurun_out/r2ae_$tool.log | tail -4
done
