#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/parity20.log 2>&1
tail -n 3 gpurun_out/parity20.log
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench20.json 2> gpurun_out/bench20.err
