#!/bin/bash
mkdir -p gpurun_out
for b in 0 18 22 29 36; do
  SSB_COST_VERBOSE=1 SSB_COST_BANDS=$b python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench17_b${b}.json 2> gpurun_out/bench17_b${b}.err
  SSB_COST_TX16=1 SSB_COST_VERBOSE=1 SSB_COST_BANDS=$b python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench17_t16_b${b}.json 2> gpurun_out/bench17_t16_b${b}.err
done
head -n 1 gpurun_out/bench17_b0.err gpurun_out/bench17_t16_b0.err
