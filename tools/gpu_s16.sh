#!/bin/bash
mkdir -p gpurun_out
for h in 000 001 071 371 370 300; do
  SSB_AGGR_HINT=$h python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench16_h${h}.json 2> gpurun_out/bench16_h${h}.err
done
SSB_AGGR_HINT=371 python -m pytest tests -m gpu -x -q -k "c1_full or small_all" 2>&1 | tail -n 2
