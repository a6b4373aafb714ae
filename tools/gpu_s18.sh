#!/bin/bash
mkdir -p gpurun_out
for b in 0 14 22; do
  SSB_COST_VERBOSE=1 SSB_COST_BANDS=$b python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_b${b}.json 2> gpurun_out/bench18_b${b}.err
done
SSB_COST_TX16=1 SSB_COST_VERBOSE=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench18_t16.json 2> gpurun_out/bench18_t16.err
head -n 1 gpurun_out/bench18_b0.err gpurun_out/bench18_t16.err
