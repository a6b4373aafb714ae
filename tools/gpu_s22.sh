#!/bin/bash
mkdir -p gpurun_out
for c in 22 32 42 33 43 23; do
  SSB_FORK_CFG=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench22_c${c}.json 2> gpurun_out/bench22_c${c}.err
done
SSB_FORK_CFG=42 timeout 120 python tools/trace_aggr.py C1 2>&1 | grep -A4 "^left\|^down"
