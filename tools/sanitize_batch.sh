#!/bin/bash
# compute-sanitizer over the batched (throughput-regime) kernels: the half-warp final pass at D = 64 / 96, the streaming
# cost kernel (32- and 16-column strips), 12-bit packed volumes
mkdir -p gpurun_out
for cfg in "small435 16" "small96 16"; do
  for tool in memcheck racecheck synccheck initcheck; do
    log=gpurun_out/sanitize_batch_${cfg%% *}_$tool.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/batch_check.py $cfg > $log 2>&1
    echo "== $cfg $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -n 1) | $(grep -E 'mismatches' $log | tail -n 1)"
  done
done
