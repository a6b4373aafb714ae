#!/bin/bash
mkdir -p gpurun_out
python tools/trace_aggr.py C1 > gpurun_out/trace5_c1.txt 2>&1
SSB_AGGR_FORK=2 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench5_fk2.json 2> gpurun_out/bench5_fk2.err
SSB_AGGR_FORK=2 python -m pytest tests -m gpu -x -q > gpurun_out/parity5.log 2>&1
tail -n 3 gpurun_out/parity5.log
grep -A8 "right+wta" gpurun_out/trace5_c1.txt
