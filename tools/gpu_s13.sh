#!/bin/bash
mkdir -p gpurun_out
for ns in 4 8 9; do
  SSB_COST_NS=$ns python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench13_ns${ns}.json 2> gpurun_out/bench13_ns${ns}.err
done
SSB_COST_NS=8 python tools/trace_aggr.py C1 2>&1 | head -6
SSB_COST_NS=8 python -m pytest tests -m gpu -x -q -k "c1_full" 2>&1 | tail -n 2
