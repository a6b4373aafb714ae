#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/parity9.log 2>&1
tail -n 3 gpurun_out/parity9.log
for e in 0 3; do
  SSB_COST_EXP=$e python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench9_exp${e}.json 2> gpurun_out/bench9_exp${e}.err
done
for tx in 24 32; do
  SSB_COST_TX=$tx python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench9_tx${tx}.json 2> gpurun_out/bench9_tx${tx}.err
done
python tools/trace_aggr.py C1 2>&1 | head -6
