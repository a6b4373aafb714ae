#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/parity23.log 2>&1
tail -n 3 gpurun_out/parity23.log
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err
