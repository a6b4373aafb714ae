#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/parity21.log 2>&1
tail -n 5 gpurun_out/parity21.log
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench21.json 2> gpurun_out/bench21.err
SSB_AGGR_NO_TMA=1 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench21_notma.json 2> gpurun_out/bench21_notma.err
timeout 120 python tools/trace_aggr.py C1 2>&1 | grep -A5 "^down\|^up" 
