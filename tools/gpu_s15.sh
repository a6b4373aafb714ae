#!/bin/bash
mkdir -p gpurun_out
for m in 4 5 6; do
  SSB_FRONT_MINB=$m python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench15_m${m}.json 2> gpurun_out/bench15_m${m}.err
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
