#!/bin/bash
# Round-1 capture 3: bench (both arms), launch list, full ncu of one frame's kernels
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 10 > gpurun_out/r01c_bench_c1.json 2> gpurun_out/r01c_bench_c1.err
python bench.py --impl reference --steps 50 --warmup 10 > gpurun_out/r01c_bench_c1_reference.json 2> gpurun_out/r01c_bench_c1_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:front7|cost_kernel|aggr_|lr_median|dilate' -s 24 -c 8 -o gpurun_out/r01c_full \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_full.log 2>&1
python tools/trace_aggr.py C1 > gpurun_out/r01c_trace.txt 2>&1
ls -la gpurun_out | tail -8
