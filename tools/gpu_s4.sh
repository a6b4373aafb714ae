#!/bin/bash
mkdir -p gpurun_out
python tools/trace_aggr.py C1 > gpurun_out/trace_c1.txt 2>&1
SSB_AGGR_FORK=2 python tools/trace_aggr.py C1 > gpurun_out/trace_c1_fork.txt 2>&1
cat gpurun_out/trace_c1.txt
