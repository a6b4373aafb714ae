#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/parity11.log 2>&1
tail -n 5 gpurun_out/parity11.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err
for tx in 24 32; do
  SSB_COST_TX=$tx python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench11_tx${tx}.json 2> gpurun_out/bench11_tx${tx}.err
done
python tools/trace_aggr.py C1 2>&1 | head -6
