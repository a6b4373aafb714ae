#!/bin/bash
# compute-sanitizer over a short mixed-mode run (small images): memcheck, racecheck (shared-memory hazards), synccheck
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck synccheck}; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/stress.py ${2:-small435} 40 > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -n 1)"
  grep -E "mismatches" gpurun_out/sanitize_$tool.log | tail -n 1
done
