#!/bin/bash
mkdir -p gpurun_out
for fk in 0 2; do
  SSB_AGGR_FORK=$fk python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench3_fk${fk}.json 2> gpurun_out/bench3_fk${fk}.err
done
SSB_AGGR_FORK=2 python -m pytest tests -m gpu -x -q > gpurun_out/parity3.log 2>&1
tail -n 3 gpurun_out/parity3.log
