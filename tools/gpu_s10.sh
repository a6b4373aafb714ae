#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:cost_kernel' -s 4 -c 1 -o gpurun_out/prof_s10 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s10.log 2>&1
tail -n 2 gpurun_out/ncu_s10.log | cut -c1-200
