#!/bin/bash
mkdir -p gpurun_out
for m in 3 4 24; do
  SSB_COST_VERBOSE=1 SSB_COST_MINB=$m python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench19_m${m}.json 2> gpurun_out/bench19_m${m}.err
  head -n 1 gpurun_out/bench19_m${m}.err
done
