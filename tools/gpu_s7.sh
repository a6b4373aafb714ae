#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/parity7.log 2>&1
tail -n 3 gpurun_out/parity7.log
for e in 0 1 3; do
  SSB_COST_EXP=$e python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_exp${e}.json 2> gpurun_out/bench7_exp${e}.err
done
for tx in 24 32; do
  SSB_COST_TX=$tx python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench7_tx${tx}.json 2> gpurun_out/bench7_tx${tx}.err
done
