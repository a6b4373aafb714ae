#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2 3; do
  SSB_COST_EXP=$e python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench6_exp${e}.json 2> gpurun_out/bench6_exp${e}.err
done
