#!/bin/bash
# GPU session 1: instruction probes, baseline bench, left||down fork experiment, source-level ncu of cost + final pass
mkdir -p gpurun_out
gpurun_out/ubench > gpurun_out/ubench.txt 2>&1
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err
SSB_AGGR_FORK=1 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_fork1.json 2> gpurun_out/bench_fork1.err
SSB_AGGR_FORK=2 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_fork2.json 2> gpurun_out/bench_fork2.err
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:aggr_wta_kernel|cost_kernel' -s 8 -c 2 -o gpurun_out/prof_s1 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_s1.log 2>&1
ls -la gpurun_out
